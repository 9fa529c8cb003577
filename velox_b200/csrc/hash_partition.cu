// Key hashing (VectorHasher::hash), partition ids (HashPartitionFunction), partition scatter order
// and gather. HBM-bound integer work: one coalesced pass per kernel, no tensor cores.
#include "common.cuh"

namespace vb2 {

constexpr int kMaxKeyCols = 8;
struct ColSet {
  vb2_column c[kMaxKeyCols];
  int n;
};

// Decodes row -> (base index, is_null) for flat / dictionary / constant columns.
__device__ __forceinline__ bool decode_row(const vb2_column& c, int64_t row, int64_t& base) {
  if (c.encoding == VB2_FLAT) {
    base = row;
    return c.nulls && !bit_at(c.nulls, row);
  }
  if (c.encoding == VB2_DICTIONARY) {
    if (c.nulls && !bit_at(c.nulls, row)) { base = 0; return true; }
    base = c.indices[row];
    return c.dict_nulls && !bit_at(c.dict_nulls, base);
  }
  base = 0;
  return c.nulls && !bit_at(c.nulls, 0);
}

__device__ __forceinline__ uint64_t hash_value(const vb2_column& c, int64_t base) {
  switch (c.type) {
    case VB2_BIGINT: return twang_mix64(static_cast<uint64_t>(reinterpret_cast<const int64_t*>(c.values)[base]));
    case VB2_INTEGER: return jenkins_rev_mix32(static_cast<uint32_t>(reinterpret_cast<const int32_t*>(c.values)[base]));
    case VB2_DOUBLE: return hash_f64(reinterpret_cast<const double*>(c.values)[base]);
    case VB2_BOOLEAN: return bit_at(reinterpret_cast<const uint64_t*>(c.values), base) ? 1 : 0;
    default: {
      const int32_t* off = reinterpret_cast<const int32_t*>(c.values);
      const int32_t b = off[base], e = off[base + 1];
      return hash_bytes(1, reinterpret_cast<const uint8_t*>(c.aux) + b, e - b);
    }
  }
}

__global__ void hash_columns_kernel(const __grid_constant__ ColSet cs, int64_t rows, uint64_t* __restrict__ out) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    uint64_t h = 0;
    for (int k = 0; k < cs.n; ++k) {
      int64_t base;
      const bool is_null = decode_row(cs.c[k], r, base);
      const uint64_t hk = is_null ? kNullHash : hash_value(cs.c[k], base);
      h = k == 0 ? hk : hash_mix(h, hk);
    }
    out[r] = h;
  }
}

__global__ void partition_ids_kernel(const uint64_t* __restrict__ hashes, int64_t rows, uint32_t parts, uint32_t* __restrict__ ids) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < rows;
       r += static_cast<int64_t>(gridDim.x) * blockDim.x)
    ids[r] = static_cast<uint32_t>(hashes[r] % parts);
}

// --- stable partition order: per-block histograms -> exclusive offsets -> scatter -------------
constexpr int kPartThreads = 256;
constexpr int kPartRowsPerBlock = 4096;
constexpr int kMaxParts = 64;

// Partition id of row r: precomputed (ids), or computed on the fly from ONE flat NULL-free integer key
// column exactly as vb2k_hash_columns + vb2k_partition_ids would (folly::hasher of the value, % parts).
struct PartSrc {
  const uint32_t* ids;
  const void* key;
  int32_t is64;
};
__device__ __forceinline__ uint32_t part_id(const PartSrc& s, int64_t r, uint32_t parts) {
  if (s.ids) return s.ids[r];
  const uint64_t h = s.is64 ? twang_mix64(static_cast<uint64_t>(reinterpret_cast<const int64_t*>(s.key)[r]))
                            : static_cast<uint64_t>(jenkins_rev_mix32(static_cast<uint32_t>(reinterpret_cast<const int32_t*>(s.key)[r])));
  return static_cast<uint32_t>(h % parts);
}

__global__ void part_hist_kernel(const __grid_constant__ PartSrc ids, int64_t rows, int parts, int32_t* __restrict__ block_hist) {
  __shared__ int32_t h[kMaxParts];
  for (int p = threadIdx.x; p < parts; p += blockDim.x) h[p] = 0;
  __syncthreads();
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kPartRowsPerBlock;
  for (int i = threadIdx.x; i < kPartRowsPerBlock; i += blockDim.x) {
    const int64_t r = r0 + i;
    if (r < rows) atomicAdd(&h[part_id(ids, r, parts)], 1);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < parts; p += blockDim.x) block_hist[static_cast<int64_t>(blockIdx.x) * parts + p] = h[p];
}

// One block of 1024 threads: for each partition, the exclusive scan of the block counts (partition-major order).
// Every thread owns a contiguous run of blocks: it sums the run, the run totals are scanned across the block
// (warp shuffles + warp totals), then the thread writes the bases of its run. (The first version walked all
// blocks with one thread per partition: fine for a filtered Q14 probe side, 50+ ms for 500 M rows.)
constexpr int kOffsetThreads = 1024;
__global__ void __launch_bounds__(kOffsetThreads) part_offsets_kernel(const int32_t* __restrict__ block_hist, int64_t nblocks, int parts,
                                                                      int64_t* __restrict__ counts, int64_t* __restrict__ block_base) {
  __shared__ int64_t warp_total[kOffsetThreads / kWarp];
  __shared__ int64_t part_total;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int64_t per = (nblocks + kOffsetThreads - 1) / kOffsetThreads;
  const int64_t b0 = t * per < nblocks ? t * per : nblocks;
  const int64_t b1 = b0 + per < nblocks ? b0 + per : nblocks;
  int64_t start = 0;  // rows of the partitions before p
  for (int p = 0; p < parts; ++p) {
    int64_t mine = 0;
    for (int64_t b = b0; b < b1; ++b) mine += block_hist[b * parts + p];
    int64_t incl = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int64_t up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) warp_total[warp] = incl;
    __syncthreads();
    if (warp == 0) {
      int64_t w = warp_total[lane];
      int64_t wi = w;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int64_t up = __shfl_up_sync(0xffffffffu, wi, o);
        if (lane >= o) wi += up;
      }
      warp_total[lane] = wi - w;  // exclusive
      if (lane == 31) part_total = wi;
    }
    __syncthreads();
    int64_t run = start + warp_total[warp] + incl - mine;
    for (int64_t b = b0; b < b1; ++b) {
      block_base[b * parts + p] = run;
      run += block_hist[b * parts + p];
    }
    const int64_t total = part_total;
    if (t == 0) counts[p] = total;
    start += total;
    __syncthreads();
  }
}

__global__ void part_scatter_kernel(const __grid_constant__ PartSrc ids, int64_t rows, int parts, const int64_t* __restrict__ block_base,
                                    int32_t* __restrict__ order) {
  // Stable within a partition: each warp ranks its rows with ballots, warps proceed in order.
  __shared__ int64_t base[kMaxParts];
  __shared__ int32_t warp_counts[kPartThreads / kWarp][kMaxParts];
  for (int p = threadIdx.x; p < parts; p += blockDim.x) base[p] = block_base[static_cast<int64_t>(blockIdx.x) * parts + p];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kPartRowsPerBlock;
  for (int i0 = 0; i0 < kPartRowsPerBlock; i0 += kPartThreads) {
    const int64_t r = r0 + i0 + threadIdx.x;
    const bool live = r < rows;
    const uint32_t id = live ? part_id(ids, r, parts) : 0xffffffffu;
    // rank among lanes of the warp with the same partition
    const unsigned peers = __match_any_sync(0xffffffffu, id);
    const int rank = __popc(peers & ((1u << lane) - 1));
    const int cnt = __popc(peers);
    for (int p = lane; p < parts; p += kWarp) warp_counts[warp][p] = 0;
    __syncwarp();
    if (live && rank == 0) warp_counts[warp][id] = cnt;
    __syncthreads();
    if (live) {
      int64_t pos = base[id];
      for (int w = 0; w < warp; ++w) pos += warp_counts[w][id];
      order[pos + rank] = static_cast<int32_t>(r);
    }
    __syncthreads();
    for (int p = threadIdx.x; p < parts; p += blockDim.x) {
      int32_t s = 0;
      for (int w = 0; w < kPartThreads / kWarp; ++w) s += warp_counts[w][p];
      base[p] += s;
    }
    __syncthreads();
  }
}

// --- fixed-capacity segments (sync-free shuffle) ---------------------------------------------------
// Same three passes as above with the partition id computed inline from a BIGINT key and rows
// written straight into their destination segment — no id / order arrays, no host round trip.
constexpr int kMaxSegCols = 4;
struct SegCols {
  const void* in[kMaxSegCols];
  void* out[kMaxSegCols];
  int32_t bytes[kMaxSegCols];
  int n;
};
__device__ __forceinline__ int64_t live_rows(int64_t rows, const int64_t* rows_dev) {
  if (!rows_dev) return rows;
  const int64_t d = *rows_dev;
  return d < rows ? d : rows;
}
__global__ void seg_hist_kernel(const int64_t* __restrict__ keys, int64_t rows, const int64_t* __restrict__ rows_dev, uint32_t parts,
                                int32_t* __restrict__ block_hist) {
  __shared__ int32_t h[kMaxParts];
  rows = live_rows(rows, rows_dev);
  for (int p = threadIdx.x; p < parts; p += blockDim.x) h[p] = 0;
  __syncthreads();
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kPartRowsPerBlock;
  for (int i = threadIdx.x; i < kPartRowsPerBlock; i += blockDim.x) {
    const int64_t r = r0 + i;
    if (r < rows) atomicAdd(&h[twang_mix64(static_cast<uint64_t>(keys[r])) % parts], 1);
  }
  __syncthreads();
  for (int p = threadIdx.x; p < parts; p += blockDim.x) block_hist[static_cast<int64_t>(blockIdx.x) * parts + p] = h[p];
}
// One block: per partition, exclusive scan of the block counts (positions inside the segment).
__global__ void seg_offsets_kernel(const int32_t* __restrict__ block_hist, int64_t nblocks, int parts, int64_t segcap,
                                   int64_t* __restrict__ counts, int64_t* __restrict__ block_base, int32_t* __restrict__ overflow) {
  const int p = threadIdx.x;
  if (p >= parts) return;
  int64_t run = 0;
  for (int64_t b = 0; b < nblocks; ++b) {
    block_base[b * parts + p] = run;
    run += block_hist[b * parts + p];
  }
  counts[p] = run;
  if (run > segcap) atomicExch(overflow, 1);
}
__global__ void seg_scatter_kernel(const int64_t* __restrict__ keys, int64_t rows, const int64_t* __restrict__ rows_dev, uint32_t parts,
                                   int64_t segcap, const int64_t* __restrict__ block_base, int64_t* __restrict__ seg_keys,
                                   const __grid_constant__ SegCols cols) {
  __shared__ int64_t base[kMaxParts];
  __shared__ int32_t warp_counts[kPartThreads / kWarp][kMaxParts];
  rows = live_rows(rows, rows_dev);
  for (int p = threadIdx.x; p < parts; p += blockDim.x) base[p] = block_base[static_cast<int64_t>(blockIdx.x) * parts + p];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t r0 = static_cast<int64_t>(blockIdx.x) * kPartRowsPerBlock;
  for (int i0 = 0; i0 < kPartRowsPerBlock; i0 += kPartThreads) {
    const int64_t r = r0 + i0 + threadIdx.x;
    const bool live = r < rows;
    const int64_t key = live ? keys[r] : 0;
    const uint32_t id = live ? static_cast<uint32_t>(twang_mix64(static_cast<uint64_t>(key)) % parts) : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, id);
    const int rank = __popc(peers & ((1u << lane) - 1));
    const int cnt = __popc(peers);
    for (int p = lane; p < parts; p += kWarp) warp_counts[warp][p] = 0;
    __syncwarp();
    if (live && rank == 0) warp_counts[warp][id] = cnt;
    __syncthreads();
    if (live) {
      int64_t pos = base[id];
      for (int w = 0; w < warp; ++w) pos += warp_counts[w][id];
      pos += rank;
      if (pos < segcap) {  // beyond: dropped, the overflow flag is already set
        const int64_t at = static_cast<int64_t>(id) * segcap + pos;
        seg_keys[at] = key;
        for (int c = 0; c < cols.n; ++c) {
          if (cols.bytes[c] == 8) reinterpret_cast<uint64_t*>(cols.out[c])[at] = reinterpret_cast<const uint64_t*>(cols.in[c])[r];
          else reinterpret_cast<uint32_t*>(cols.out[c])[at] = reinterpret_cast<const uint32_t*>(cols.in[c])[r];
        }
      }
    }
    __syncthreads();
    for (int p = threadIdx.x; p < parts; p += blockDim.x) {
      int32_t s = 0;
      for (int w = 0; w < kPartThreads / kWarp; ++w) s += warp_counts[w][p];
      base[p] += s;
    }
    __syncthreads();
  }
}

// flag <- 1 if any key other than the sentinel lies outside [lo, hi]
__global__ void key_range_check_kernel(const int64_t* __restrict__ keys, int64_t n, int64_t lo, int64_t hi, int32_t* __restrict__ flag) {
  bool bad = false;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t k = keys[i];
    bad |= k != VB2_SENTINEL_KEY && (k < lo || k > hi);
  }
  if (__any_sync(0xffffffffu, bad) && (threadIdx.x & 31) == 0) atomicExch(flag, 1);
}

template <class T>
__global__ void gather_kernel(const T* __restrict__ in, const int32_t* __restrict__ order, int64_t n, T* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = in[order[i]];
}

static unsigned grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace vb2

using namespace vb2;

extern "C" {

int vb2k_hash_columns(const vb2_column* cols, int32_t ncols, int64_t rows, uint64_t* hashes, void* stream) {
  if (ncols < 1 || ncols > kMaxKeyCols) return fail_msg(VB2_ERR_INVALID, "hash_columns: 1..8 key columns");
  if (rows <= 0) return VB2_OK;
  ColSet cs;
  cs.n = ncols;
  for (int i = 0; i < ncols; ++i) cs.c[i] = cols[i];
  hash_columns_kernel<<<vb2::counted(grid_for(rows, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(cs, rows, hashes);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_partition_ids(const uint64_t* hashes, int64_t rows, int32_t num_partitions, uint32_t* ids, void* stream) {
  if (num_partitions < 1) return fail_msg(VB2_ERR_INVALID, "partition_ids: num_partitions < 1");
  if (rows <= 0) return VB2_OK;
  partition_ids_kernel<<<vb2::counted(grid_for(rows, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(hashes, rows, static_cast<uint32_t>(num_partitions), ids);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

static int partition_order(const PartSrc& src, int64_t rows, int32_t num_partitions, int64_t* counts, int32_t* row_order, void* stream);

int vb2k_partition_scatter_order(const uint32_t* ids, int64_t rows, int32_t num_partitions, int64_t* counts,
                                 int32_t* row_order, void* stream) {
  return partition_order(PartSrc{ids, nullptr, 0}, rows, num_partitions, counts, row_order, stream);
}

int vb2k_partition_order_key(const void* key_values, int32_t key_is64, int64_t rows, int32_t num_partitions, int64_t* counts, int32_t* row_order,
                             void* stream) {
  if (!key_values && rows > 0) return fail_msg(VB2_ERR_INVALID, "partition_order_key: no key column");
  return partition_order(PartSrc{nullptr, key_values, key_is64}, rows, num_partitions, counts, row_order, stream);
}

static int partition_order(const PartSrc& ids, int64_t rows, int32_t num_partitions, int64_t* counts, int32_t* row_order, void* stream) {
  if (num_partitions < 1 || num_partitions > kMaxParts) return fail_msg(VB2_ERR_INVALID, "partition_scatter_order: 1..64 partitions");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows <= 0) {
    VB2_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * num_partitions, st));
    return VB2_OK;
  }
  const int64_t nblocks = (rows + kPartRowsPerBlock - 1) / kPartRowsPerBlock;
  int32_t* hist = nullptr;
  int64_t* base = nullptr;
  VB2_CUDA_OK(cudaMallocAsync(&hist, sizeof(int32_t) * nblocks * num_partitions, st));
  VB2_CUDA_OK(cudaMallocAsync(&base, sizeof(int64_t) * nblocks * num_partitions, st));
  part_hist_kernel<<<vb2::counted(static_cast<unsigned>(nblocks)), kPartThreads, 0, st>>>(ids, rows, num_partitions, hist);
  part_offsets_kernel<<<vb2::counted(1), kOffsetThreads, 0, st>>>(hist, nblocks, num_partitions, counts, base);
  part_scatter_kernel<<<vb2::counted(static_cast<unsigned>(nblocks)), kPartThreads, 0, st>>>(ids, rows, num_partitions, base, row_order);
  VB2_CUDA_OK(cudaGetLastError());
  VB2_CUDA_OK(cudaFreeAsync(hist, st));
  VB2_CUDA_OK(cudaFreeAsync(base, st));
  return VB2_OK;
}

int vb2k_partition_segments(const int64_t* keys, const void* const* cols, const int32_t* col_elem_bytes, int32_t ncols, int64_t rows,
                            const int64_t* rows_dev, int32_t num_partitions, int64_t segcap, int64_t* seg_keys, void* const* seg_cols,
                            int64_t* counts, int32_t* overflow, void* stream) {
  if (num_partitions < 1 || num_partitions > kMaxParts) return fail_msg(VB2_ERR_INVALID, "partition_segments: 1..64 partitions");
  if (ncols < 0 || ncols > kMaxSegCols) return fail_msg(VB2_ERR_UNSUPPORTED, "partition_segments: at most 4 payload columns");
  if (segcap <= 0) return fail_msg(VB2_ERR_INVALID, "partition_segments: segment capacity must be positive");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SegCols sc;
  sc.n = ncols;
  for (int c = 0; c < ncols; ++c) {
    if (col_elem_bytes[c] != 4 && col_elem_bytes[c] != 8) return fail_msg(VB2_ERR_INVALID, "partition_segments: 4- or 8-byte payload columns");
    sc.in[c] = cols[c];
    sc.out[c] = seg_cols[c];
    sc.bytes[c] = col_elem_bytes[c];
  }
  // every key slot starts as the sentinel (bytes 0x80): segment tails stay that way
  VB2_CUDA_OK(cudaMemsetAsync(seg_keys, 0x80, sizeof(int64_t) * static_cast<size_t>(num_partitions) * segcap, st));
  if (rows <= 0) {
    VB2_CUDA_OK(cudaMemsetAsync(counts, 0, sizeof(int64_t) * num_partitions, st));
    return VB2_OK;
  }
  const int64_t nblocks = (rows + kPartRowsPerBlock - 1) / kPartRowsPerBlock;
  int32_t* hist = nullptr;
  int64_t* base = nullptr;
  VB2_CUDA_OK(cudaMallocAsync(&hist, sizeof(int32_t) * nblocks * num_partitions, st));
  VB2_CUDA_OK(cudaMallocAsync(&base, sizeof(int64_t) * nblocks * num_partitions, st));
  seg_hist_kernel<<<vb2::counted(static_cast<unsigned>(nblocks)), kPartThreads, 0, st>>>(keys, rows, rows_dev, static_cast<uint32_t>(num_partitions), hist);
  seg_offsets_kernel<<<vb2::counted(1), kMaxParts, 0, st>>>(hist, nblocks, num_partitions, segcap, counts, base, overflow);
  seg_scatter_kernel<<<vb2::counted(static_cast<unsigned>(nblocks)), kPartThreads, 0, st>>>(keys, rows, rows_dev, static_cast<uint32_t>(num_partitions), segcap, base,
                                                                              seg_keys, sc);
  VB2_CUDA_OK(cudaGetLastError());
  VB2_CUDA_OK(cudaFreeAsync(hist, st));
  VB2_CUDA_OK(cudaFreeAsync(base, st));
  return VB2_OK;
}

int vb2k_key_range_check(const int64_t* keys, int64_t n, int64_t lo, int64_t hi, int32_t* flag, void* stream) {
  if (n <= 0) return VB2_OK;
  key_range_check_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(keys, n, lo, hi, flag);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_gather(const void* in, const int32_t* order, int64_t n, int32_t elem_bytes, void* out, void* stream) {
  if (n <= 0) return VB2_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 8)
    gather_kernel<uint64_t><<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(reinterpret_cast<const uint64_t*>(in), order, n, reinterpret_cast<uint64_t*>(out));
  else if (elem_bytes == 4)
    gather_kernel<uint32_t><<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(in), order, n, reinterpret_cast<uint32_t*>(out));
  else if (elem_bytes == 1)
    gather_kernel<uint8_t><<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(reinterpret_cast<const uint8_t*>(in), order, n, reinterpret_cast<uint8_t*>(out));
  else
    return fail_msg(VB2_ERR_INVALID, "gather: elem_bytes must be 1, 4 or 8");
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

}  // extern "C"
