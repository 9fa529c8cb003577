// Hash join build / probe and the exclusive scan that sizes the probe output.
//
// B200-native take on exec::HashTable for joins:
//   * array mode when the build key range is dense enough (slot = key - min): for TPC-H Q14 the
//     20 M-entry int32 head[] array (80 MB) stays resident in the 126 MB L2, so a probe is one
//     L2 hit instead of the reference's tag + row-pointer chase (exec/HashTable.cpp:610-725);
//   * hash mode otherwise: uint64 keys[capacity] claimed by CAS + head[]; duplicates are
//     chained through next[] exactly like the reference's next-row pointer
//     (exec/HashTable.cpp:1518 insertForJoin), pushed with atomicExch;
//   * probe emits (probe row, build row) pairs in probe-row order via count -> scan -> emit
//     (listJoinResults, exec/HashTable.cpp:2133-2350); rows with NULL keys never match.
#include "common.cuh"

namespace vb2 {

__device__ __forceinline__ int64_t find_slot(const vb2_join_table& t, uint64_t key, bool insert) {
  if (t.mode == 0) {
    const int64_t s = static_cast<int64_t>(key) - t.key_min;
    return (s >= 0 && s < t.capacity) ? s : -1;
  }
  const uint64_t mask = static_cast<uint64_t>(t.capacity - 1);
  uint64_t slot = twang_mix64(key) & mask;
  for (uint64_t probes = 0; probes <= mask; ++probes) {
    uint64_t cur = t.keys[slot];
    if (cur == VB2_EMPTY_KEY) {
      if (!insert) return -1;
      cur = atomicCAS(reinterpret_cast<unsigned long long*>(t.keys + slot), VB2_EMPTY_KEY, static_cast<unsigned long long>(key));
      if (cur == VB2_EMPTY_KEY) return static_cast<int64_t>(slot);
    }
    if (cur == key) return static_cast<int64_t>(slot);
    slot = (slot + 1) & mask;
  }
  return -2;  // full
}

__global__ void join_build_kernel(const __grid_constant__ vb2_join_table t, const uint64_t* __restrict__ keys,
                                  const uint64_t* __restrict__ valid, int64_t n, int32_t* __restrict__ error_flag) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < n; r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (valid && !bit_at(valid, r)) continue;  // NULL keys are not inserted (exec/HashBuild.cpp:475-479)
    const int64_t slot = find_slot(t, keys[r], true);
    if (slot < 0) { atomicCAS(error_flag, 0, slot == -2 ? 100 : 101); continue; }
    // push at the head of the chain; next[] stores row + 1
    const int32_t prev = atomicExch(t.head + slot, static_cast<int32_t>(r + 1));
    t.next[r] = prev;
    if (prev != 0) atomicCAS(error_flag + 1, 0, 1);  // duplicate build keys present (informational)
  }
}

__global__ void join_probe_count_kernel(const __grid_constant__ vb2_join_table t, const uint64_t* __restrict__ keys,
                                        const uint64_t* __restrict__ valid, int64_t n, int32_t* __restrict__ counts) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < n; r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int32_t c = 0;
    if (!valid || bit_at(valid, r)) {
      const int64_t slot = find_slot(t, keys[r], false);
      if (slot >= 0)
        for (int32_t m = t.head[slot]; m != 0; m = t.next[m - 1]) ++c;
    }
    counts[r] = c;
  }
}

// Unique build keys (no chains): one probe per row. The match flags of a warp become one bitmap word
// through __ballot_sync (the compaction of the matches is then the same ordered bitmap expansion the
// filter uses), the matched build row of every probe row is kept beside it.
__global__ void join_probe_unique_kernel(const __grid_constant__ vb2_join_table t, const uint64_t* __restrict__ keys,
                                         const uint64_t* __restrict__ valid, int64_t n, uint32_t* __restrict__ hit_bits,
                                         int32_t* __restrict__ hits) {
  const int64_t nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t r = (w << 5) + lane;
    int32_t m = 0;
    if (r < n && (!valid || bit_at(valid, r))) {
      const int64_t slot = find_slot(t, keys[r], false);
      if (slot >= 0) m = t.head[slot];
    }
    if (r < n) hits[r] = m - 1;
    const unsigned word = __ballot_sync(0xffffffffu, m != 0);
    if (lane == 0) hit_bits[w] = word;
  }
}

__global__ void join_probe_emit_kernel(const __grid_constant__ vb2_join_table t, const uint64_t* __restrict__ keys,
                                       const uint64_t* __restrict__ valid, int64_t n, const int64_t* __restrict__ offsets,
                                       int32_t* __restrict__ probe_rows, int32_t* __restrict__ build_rows) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < n; r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (valid && !bit_at(valid, r)) continue;
    const int64_t slot = find_slot(t, keys[r], false);
    if (slot < 0) continue;
    int64_t pos = offsets[r];
    for (int32_t m = t.head[slot]; m != 0; m = t.next[m - 1]) {
      probe_rows[pos] = static_cast<int32_t>(r);
      build_rows[pos] = m - 1;
      ++pos;
    }
  }
}

// ---- exclusive scan of int32 counts into int64 offsets ----------------------------------------
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;  // per thread -> 2048 per block

__global__ void scan_block_sums_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ block_sums) {
  __shared__ int64_t ws[kScanThreads / kWarp];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanThreads * kScanItems;
  int64_t s = 0;
  for (int j = 0; j < kScanItems; ++j) {
    const int64_t i = base + static_cast<int64_t>(j) * kScanThreads + threadIdx.x;
    if (i < n) s += in[i];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    int64_t t = 0;
    for (int w = 0; w < kScanThreads / kWarp; ++w) t += ws[w];
    block_sums[blockIdx.x] = t;
  }
}
__global__ void scan_offsets_kernel(int64_t* __restrict__ block_sums, int64_t nblocks, int64_t* __restrict__ total) {
  __shared__ int64_t carry;
  __shared__ int64_t tmp[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < nblocks; b0 += 1024) {
    const int64_t b = b0 + threadIdx.x;
    const int64_t v = b < nblocks ? block_sums[b] : 0;
    tmp[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int64_t t = threadIdx.x >= o ? tmp[threadIdx.x - o] : 0;
      __syncthreads();
      tmp[threadIdx.x] += t;
      __syncthreads();
    }
    if (b < nblocks) block_sums[b] = carry + tmp[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += tmp[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void scan_write_kernel(const int32_t* __restrict__ in, int64_t n, const int64_t* __restrict__ block_offsets,
                                  int64_t* __restrict__ out) {
  // thread t owns kScanItems consecutive elements
  __shared__ int64_t ws[kScanThreads / kWarp];
  const int64_t base = static_cast<int64_t>(blockIdx.x) * kScanThreads * kScanItems + static_cast<int64_t>(threadIdx.x) * kScanItems;
  int32_t v[kScanItems];
  int64_t run = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    v[j] = base + j < n ? in[base + j] : 0;
    run += v[j];
  }
  int64_t incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int64_t x = __shfl_up_sync(0xffffffffu, incl, o);
    if ((threadIdx.x & 31) >= o) incl += x;
  }
  if ((threadIdx.x & 31) == 31) ws[threadIdx.x >> 5] = incl;
  __syncthreads();
  int64_t pre = block_offsets[blockIdx.x];
  for (int w = 0; w < (threadIdx.x >> 5); ++w) pre += ws[w];
  int64_t excl = pre + incl - run;
#pragma unroll
  for (int j = 0; j < kScanItems; ++j) {
    if (base + j < n) out[base + j] = excl;
    excl += v[j];
  }
}

static unsigned grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace vb2

using namespace vb2;

extern "C" {

int vb2k_join_build(const vb2_join_table* t, const uint64_t* build_keys, const uint64_t* valid, int64_t n,
                    int32_t* error_flag, void* stream) {
  if (!t || (t->mode == 1 && (t->capacity & (t->capacity - 1)))) return fail_msg(VB2_ERR_INVALID, "join_build: bad table");
  if (n <= 0) return VB2_OK;
  join_build_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*t, build_keys, valid, n, error_flag);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_join_probe_count(const vb2_join_table* t, const uint64_t* probe_keys, const uint64_t* valid, int64_t n,
                          int32_t* hit_counts, void* stream) {
  if (n <= 0) return VB2_OK;
  join_probe_count_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*t, probe_keys, valid, n, hit_counts);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_join_probe_unique(const vb2_join_table* t, const uint64_t* probe_keys, const uint64_t* valid, int64_t n, uint64_t* hit_bits,
                           int32_t* hits, void* stream) {
  if (n <= 0) return VB2_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // the 64-bit consumers of the bitmap read whole words: clear the tail
  VB2_CUDA_OK(cudaMemsetAsync(hit_bits + ((n + 63) >> 6) - 1, 0, sizeof(uint64_t), st));
  join_probe_unique_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(*t, probe_keys, valid, n, reinterpret_cast<uint32_t*>(hit_bits), hits);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

size_t vb2k_scan_workspace(int64_t n) {
  const int64_t nblocks = (n + kScanThreads * kScanItems - 1) / (kScanThreads * kScanItems);
  return static_cast<size_t>(nblocks < 1 ? 1 : nblocks) * sizeof(int64_t);
}

int vb2k_exclusive_scan_i32(const int32_t* in, int64_t n, int64_t* out, int64_t* total_out, void* workspace,
                            size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n <= 0) {
    VB2_CUDA_OK(cudaMemsetAsync(total_out, 0, sizeof(int64_t), st));
    return VB2_OK;
  }
  if (workspace_bytes < vb2k_scan_workspace(n)) return fail_msg(VB2_ERR_INVALID, "exclusive_scan: workspace too small");
  const int64_t nblocks = (n + kScanThreads * kScanItems - 1) / (kScanThreads * kScanItems);
  int64_t* sums = reinterpret_cast<int64_t*>(workspace);
  scan_block_sums_kernel<<<vb2::counted(static_cast<unsigned>(nblocks)), kScanThreads, 0, st>>>(in, n, sums);
  scan_offsets_kernel<<<vb2::counted(1), 1024, 0, st>>>(sums, nblocks, total_out);
  scan_write_kernel<<<vb2::counted(static_cast<unsigned>(nblocks)), kScanThreads, 0, st>>>(in, n, sums, out);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_join_probe_emit(const vb2_join_table* t, const uint64_t* probe_keys, const uint64_t* valid, int64_t n,
                         const int64_t* offsets, int32_t* probe_rows, int32_t* build_rows, void* stream) {
  if (n <= 0) return VB2_OK;
  join_probe_emit_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*t, probe_keys, valid, n, offsets, probe_rows, build_rows);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

}  // extern "C"
