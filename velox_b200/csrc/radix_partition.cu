// Radix partitioning of a batch by the TOP bits of its group-table hash, in front of a
// high-cardinality aggregation ("hard part (ii)").
//
// A hash-mode group table places key k at slot twang_mix64(k) >> (64 - log2(capacity)): the top
// bits of the hash pick the slot, so rows ordered by the top 8 bits of their hash visit the table
// slice by slice — 1/256 of the table at a time, which stays resident in the 126 MB L2 while its
// rows are folded in. The random 32-byte read-modify-write per input row (198 B of DRAM traffic
// per row at 100 M groups, profiles/r01_ncu_config5_group_update.json) becomes an L2 hit; DRAM sees
// the streaming passes of the partitioner and each table slice once per batch.
// The reference hides the same latency on CPUs by interleaving four probes and prefetching
// (exec/HashTable.cpp:485-519); partitioning is the throughput-machine answer.
//
//   radix_hist     per-chunk histogram over the 256 partitions + a HyperLogLog sketch of the keys
//                  (distinct-count estimate: sizes the table before any row is inserted)
//   radix_offsets  exclusive scan: where every chunk's run of every partition starts
//   radix_scatter  keys and up to 4 fixed-width payload columns into partition order
#include "common.cuh"

namespace vb2 {

constexpr int kRadixParts = 256;
constexpr int kRadixThreads = 256;
constexpr int kRadixChunkRows = 8192;
constexpr int kHllBits = 12;  // 4096 registers: standard error 1.6 %
constexpr int kRadixMaxCols = 4;

struct RadixKey {
  const uint64_t* norm;   // normalized keys, or NULL: the single flat integer key column below
  const void* values;
  int32_t is64;
  int64_t min;            // normalized key = v - min + 1 (vb2k_normalize_keys with one column)
};
__device__ __forceinline__ uint64_t radix_key(const RadixKey& k, int64_t r) {
  if (k.norm) return k.norm[r];
  const int64_t v = k.is64 ? reinterpret_cast<const int64_t*>(k.values)[r] : reinterpret_cast<const int32_t*>(k.values)[r];
  return static_cast<uint64_t>(v - k.min) + 1;
}

__global__ void __launch_bounds__(kRadixThreads) radix_hist_kernel(const __grid_constant__ RadixKey key, int64_t n, int64_t nchunks,
                                                                   int32_t* __restrict__ chunk_hist, int32_t* __restrict__ hll) {
  __shared__ int32_t h[kRadixParts];
  __shared__ int32_t regs[1 << kHllBits];
  for (int i = threadIdx.x; i < (1 << kHllBits); i += kRadixThreads) regs[i] = 0;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    h[threadIdx.x] = 0;
    __syncthreads();
    const int64_t r0 = c * kRadixChunkRows;
    for (int i = threadIdx.x; i < kRadixChunkRows; i += kRadixThreads) {
      const int64_t r = r0 + i;
      if (r < n) {
        const uint64_t hash = twang_mix64(radix_key(key, r));
        atomicAdd(&h[hash >> 56], 1);
        // HyperLogLog over the LOW bits (the top bits are the partition): 1/8 of the rows are enough
        if ((hash & 7u) == 0) {
          const uint32_t idx = static_cast<uint32_t>(hash >> 3) & ((1u << kHllBits) - 1u);
          const uint64_t rest = (hash >> (3 + kHllBits)) | (1ull << (56 - 3 - kHllBits));  // bits below the partition bits, with a stop bit
          const int rho = __ffsll(static_cast<long long>(rest));  // position of the first set bit, 1-based
          atomicMax(&regs[idx], rho);
        }
      }
    }
    __syncthreads();
    chunk_hist[c * kRadixParts + threadIdx.x] = h[threadIdx.x];
    __syncthreads();
  }
  __syncthreads();
  for (int i = threadIdx.x; i < (1 << kHllBits); i += kRadixThreads)
    if (regs[i]) atomicMax(&hll[i], regs[i]);
}

// Exclusive scan of the chunk histograms per partition, in two levels so that no thread walks all
// chunks alone: kOffsetSegs segments of chunks per partition are summed in parallel (A), one block
// turns the (partition, segment) sums into starting offsets (B), every segment then scans its own
// chunks from its start (C). chunk_base[c][p] = where chunk c's run of partition p begins.
constexpr int kOffsetSegs = 64;
__global__ void __launch_bounds__(kRadixParts) radix_offsets_a_kernel(const int32_t* __restrict__ chunk_hist, int64_t nchunks, int64_t* __restrict__ seg_sums) {
  const int p = threadIdx.x, seg = blockIdx.x;
  const int64_t per = (nchunks + kOffsetSegs - 1) / kOffsetSegs;
  const int64_t c0 = seg * per, c1 = c0 + per < nchunks ? c0 + per : nchunks;
  int64_t run = 0;
  for (int64_t c = c0; c < c1; ++c) run += chunk_hist[c * kRadixParts + p];
  seg_sums[seg * kRadixParts + p] = run;
}
__global__ void __launch_bounds__(kRadixParts) radix_offsets_b_kernel(int64_t* __restrict__ seg_sums, int64_t* __restrict__ part_start) {
  __shared__ int64_t totals[kRadixParts];
  const int p = threadIdx.x;
  int64_t run = 0;
  for (int seg = 0; seg < kOffsetSegs; ++seg) {
    const int64_t v = seg_sums[seg * kRadixParts + p];
    seg_sums[seg * kRadixParts + p] = run;  // rows of p in earlier segments
    run += v;
  }
  totals[p] = run;
  __syncthreads();
  int64_t start = 0;
  for (int q = 0; q < p; ++q) start += totals[q];
  part_start[p] = start;
  if (p == kRadixParts - 1) part_start[kRadixParts] = start + run;
  for (int seg = 0; seg < kOffsetSegs; ++seg) seg_sums[seg * kRadixParts + p] += start;
}
__global__ void __launch_bounds__(kRadixParts) radix_offsets_c_kernel(const int32_t* __restrict__ chunk_hist, int64_t nchunks, const int64_t* __restrict__ seg_sums,
                                                                      int64_t* __restrict__ chunk_base) {
  const int p = threadIdx.x, seg = blockIdx.x;
  const int64_t per = (nchunks + kOffsetSegs - 1) / kOffsetSegs;
  const int64_t c0 = seg * per, c1 = c0 + per < nchunks ? c0 + per : nchunks;
  int64_t run = seg_sums[seg * kRadixParts + p];
  for (int64_t c = c0; c < c1; ++c) {
    chunk_base[c * kRadixParts + p] = run;
    run += chunk_hist[c * kRadixParts + p];
  }
}

struct RadixCols {
  const void* in[kRadixMaxCols];
  void* out[kRadixMaxCols];
  int32_t bytes[kRadixMaxCols];
  int n;
};
// One chunk per block iteration, staged through shared memory: the chunk's keys are counting-sorted
// by partition inside the block (positions from shared-memory cursors), then written out by
// consecutive threads — each partition's run of the chunk is contiguous in the output, so the stores
// are full sectors (a direct scatter wrote 8 of every 32 bytes per sector and made DRAM read and
// write every sector twice: 15.4 GB of traffic for 8 GB of data, profiles/). Payload columns follow
// through the staged source row numbers (reads stay inside the chunk: L1 / L2 hits).
__global__ void __launch_bounds__(kRadixThreads) radix_scatter_kernel(const __grid_constant__ RadixKey key, int64_t n, int64_t nchunks,
                                                                      const int64_t* __restrict__ chunk_base, uint64_t* __restrict__ out_keys,
                                                                      const __grid_constant__ RadixCols cols) {
  __shared__ uint8_t rpid[kRadixChunkRows];    // partition of every row of the chunk
  __shared__ uint16_t ssrc[kRadixChunkRows];   // rows of the chunk in partition order
  __shared__ int32_t lcount[kRadixParts], lstart[kRadixParts], lcursor[kRadixParts];
  __shared__ long long gbase[kRadixParts];
  __shared__ int wtot[kRadixThreads / kWarp];
  constexpr int kPer = kRadixChunkRows / kRadixThreads;
  for (int64_t c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const int64_t r0 = c * kRadixChunkRows;
    gbase[threadIdx.x] = chunk_base[c * kRadixParts + threadIdx.x];
    lcount[threadIdx.x] = 0;
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < kPer; ++j) {
      const int i = j * kRadixThreads + threadIdx.x;
      const int64_t r = r0 + i;
      if (r < n) {
        const uint8_t p = static_cast<uint8_t>(twang_mix64(radix_key(key, r)) >> 56);
        rpid[i] = p;
        atomicAdd(&lcount[p], 1);
      }
    }
    __syncthreads();
    {
      // exclusive prefix of the 256 counts: warp scans + warp totals
      const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
      const int v = lcount[threadIdx.x];
      int incl = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) wtot[warp] = incl;
      __syncthreads();
      int before = 0;
      for (int w = 0; w < warp; ++w) before += wtot[w];
      lstart[threadIdx.x] = before + incl - v;
      lcursor[threadIdx.x] = before + incl - v;
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < kPer; ++j) {
      const int i = j * kRadixThreads + threadIdx.x;
      if (r0 + i < n) ssrc[atomicAdd(&lcursor[rpid[i]], 1)] = static_cast<uint16_t>(i);
    }
    __syncthreads();
    const int live = n - r0 < kRadixChunkRows ? static_cast<int>(n - r0) : kRadixChunkRows;
    for (int j = threadIdx.x; j < live; j += kRadixThreads) {
      const int i = ssrc[j];
      const int p = rpid[i];
      const int64_t pos = gbase[p] + (j - lstart[p]);
      const int64_t r = r0 + i;
      out_keys[pos] = radix_key(key, r);  // second read of the chunk's keys: L1 / L2 resident
      for (int q = 0; q < cols.n; ++q) {
        if (cols.bytes[q] == 8) reinterpret_cast<uint64_t*>(cols.out[q])[pos] = reinterpret_cast<const uint64_t*>(cols.in[q])[r];
        else reinterpret_cast<uint32_t*>(cols.out[q])[pos] = reinterpret_cast<const uint32_t*>(cols.in[q])[r];
      }
    }
    __syncthreads();
  }
}

}  // namespace vb2

using namespace vb2;

extern "C" {

size_t vb2k_radix_workspace_bytes(int64_t rows) {
  const int64_t nchunks = (rows + kRadixChunkRows - 1) / kRadixChunkRows;
  // chunk histograms (int32) + chunk bases (int64) + partition starts (int64[257]) + HLL registers
  return static_cast<size_t>(nchunks) * kRadixParts * 12 + static_cast<size_t>(kOffsetSegs) * kRadixParts * 8 + (kRadixParts + 1) * 8 + (1 << kHllBits) * 4 + 512;
}

int32_t vb2k_radix_hll_registers(void) { return 1 << kHllBits; }

// Phase 1: histogram + distinct-count sketch. hll_out: device int32[vb2k_radix_hll_registers()], zeroed here.
int vb2k_radix_histogram(const uint64_t* norm_keys, const void* key_values, int32_t key_is64, int64_t key_min, int64_t rows, void* workspace,
                         size_t workspace_bytes, int32_t* hll_out, void* stream) {
  if (rows <= 0) return VB2_OK;
  if (workspace_bytes < vb2k_radix_workspace_bytes(rows)) return fail_msg(VB2_ERR_INVALID, "radix_histogram: workspace too small");
  const int64_t nchunks = (rows + kRadixChunkRows - 1) / kRadixChunkRows;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  RadixKey k{norm_keys, key_values, key_is64, key_min};
  VB2_CUDA_OK(cudaMemsetAsync(hll_out, 0, (1 << kHllBits) * 4, st));
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 6;
  radix_hist_kernel<<<vb2::counted(static_cast<unsigned>(nchunks < cap ? nchunks : cap)), kRadixThreads, 0, st>>>(k, rows, nchunks, reinterpret_cast<int32_t*>(workspace), hll_out);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

// Phase 2: keys and payload columns into partition order. part_start_out: device int64[257].
int vb2k_radix_scatter(const uint64_t* norm_keys, const void* key_values, int32_t key_is64, int64_t key_min, int64_t rows, void* workspace,
                       size_t workspace_bytes, const void* const* cols, void* const* cols_out, const int32_t* col_bytes, int32_t ncols,
                       uint64_t* keys_out, int64_t* part_start_out, void* stream) {
  if (rows <= 0) return VB2_OK;
  if (ncols < 0 || ncols > kRadixMaxCols) return fail_msg(VB2_ERR_UNSUPPORTED, "radix_scatter: at most 4 payload columns");
  if (workspace_bytes < vb2k_radix_workspace_bytes(rows)) return fail_msg(VB2_ERR_INVALID, "radix_scatter: workspace too small");
  const int64_t nchunks = (rows + kRadixChunkRows - 1) / kRadixChunkRows;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int32_t* hist = reinterpret_cast<int32_t*>(workspace);
  int64_t* base = reinterpret_cast<int64_t*>(reinterpret_cast<char*>(workspace) + (static_cast<size_t>(nchunks) * kRadixParts * 4 + 255) / 256 * 256);
  RadixKey k{norm_keys, key_values, key_is64, key_min};
  RadixCols c{};
  c.n = ncols;
  for (int i = 0; i < ncols; ++i) {
    if (col_bytes[i] != 4 && col_bytes[i] != 8) return fail_msg(VB2_ERR_INVALID, "radix_scatter: payload widths 4 or 8");
    c.in[i] = cols[i];
    c.out[i] = cols_out[i];
    c.bytes[i] = col_bytes[i];
  }
  int64_t* seg_sums = base + nchunks * kRadixParts;  // kOffsetSegs x 256 partial sums behind the chunk bases
  radix_offsets_a_kernel<<<vb2::counted(kOffsetSegs), kRadixParts, 0, st>>>(hist, nchunks, seg_sums);
  radix_offsets_b_kernel<<<vb2::counted(1), kRadixParts, 0, st>>>(seg_sums, part_start_out);
  radix_offsets_c_kernel<<<vb2::counted(kOffsetSegs), kRadixParts, 0, st>>>(hist, nchunks, seg_sums, base);
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 6;
  radix_scatter_kernel<<<vb2::counted(static_cast<unsigned>(nchunks < cap ? nchunks : cap)), kRadixThreads, 0, st>>>(k, rows, nchunks, base, keys_out, c);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

}  // extern "C"
