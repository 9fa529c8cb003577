// Generic aggregation kernels: normalized keys, find-or-insert group table, accumulator updates.
//
// B200-native take on exec::HashTable + RowContainer for GROUP BY:
//   * keys of a row are packed into ONE 64-bit normalized key from per-column value ids
//     (id = v - min + 1, 0 = NULL; exec/VectorHasher.h:523-585) — the reference's kNormalizedKey
//     mode (exec/HashTable.cpp:523) is the only mode needed because ranges come from a device
//     min/max pass over the whole batch, not from 1K-row increments;
//   * the table is SoA: uint64 keys[capacity] (open addressing, linear probing, twang_mix64),
//     accumulators are separate dense arrays indexed by slot — no row-wise RowContainer, so an
//     accumulator update is one 8-byte atomic to a 32-byte sector instead of a row RMW;
//   * null keys form a group (value id 0), as GroupingSet does for non-ignoreNullKeys tables
//     (exec/GroupingSet.cpp:448-455).
#include "common.cuh"

namespace vb2 {

constexpr int kMaxNormCols = 4;
struct NormArgs {
  vb2_column c[kMaxNormCols];
  int64_t mins[kMaxNormCols];
  uint64_t mults[kMaxNormCols];
  uint64_t ranges[kMaxNormCols];
  int n;
  int nulls_invalid;  // a NULL key column clears the row's valid bit (joins) instead of using id 0 (group by)
  int check_ranges;   // ids outside [1, range) clear the valid bit (probe side of a join)
};

__device__ __forceinline__ bool decode_row2(const vb2_column& c, int64_t row, int64_t& base) {
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

// Integer view of a key value. DOUBLE keys use canonical bits (NaN == NaN, +0 == -0) and VARCHAR
// keys their dictionary index / short-string packing, prepared by the host layer as BIGINT columns.
__device__ __forceinline__ int64_t key_value(const vb2_column& c, int64_t base) {
  switch (c.type) {
    case VB2_BIGINT: return reinterpret_cast<const int64_t*>(c.values)[base];
    case VB2_INTEGER: return reinterpret_cast<const int32_t*>(c.values)[base];
    case VB2_BOOLEAN: return bit_at(reinterpret_cast<const uint64_t*>(c.values), base) ? 1 : 0;
    default: return 0;
  }
}

__global__ void normalize_keys_kernel(const __grid_constant__ NormArgs a, const int32_t* __restrict__ sel, int64_t n,
                                      uint64_t* __restrict__ out, uint32_t* __restrict__ valid_out) {
  const int64_t nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t i = (w << 5) + lane;
    bool valid = i < n;
    if (i < n) {
      const int64_t row = sel ? sel[i] : i;
      uint64_t key = 0;
      for (int k = 0; k < a.n; ++k) {
        int64_t base;
        const bool is_null = decode_row2(a.c[k], row, base);
        uint64_t id = 0;
        if (is_null) {
          valid = valid && !a.nulls_invalid;
        } else {
          id = static_cast<uint64_t>(key_value(a.c[k], base) - a.mins[k]) + 1;
          if (a.check_ranges && (id == 0 || id >= a.ranges[k])) { valid = false; id = 0; }
        }
        key += id * a.mults[k];
      }
      out[i] = key;
    }
    if (valid_out) {
      const unsigned word = __ballot_sync(0xffffffffu, valid);
      if (lane == 0) valid_out[w] = word;
    }
  }
}

// Keys of occupied slots back to per-column values: id_k = (key / mult_k) % range_k.
__global__ void denormalize_keys_kernel(const uint64_t* __restrict__ table_keys, const int32_t* __restrict__ slots, int64_t n,
                                        int64_t min, uint64_t mult, uint64_t range, int32_t null_reserved, int32_t type,
                                        void* __restrict__ values, uint32_t* __restrict__ valid_words) {
  const int64_t nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t i = (w << 5) + lane;
    bool valid = false;
    if (i < n) {
      const uint64_t key = table_keys ? table_keys[slots[i]] : static_cast<uint64_t>(slots[i]);
      const uint64_t id = (key / mult) % range;
      valid = !(null_reserved && id == 0);
      const int64_t v = valid ? static_cast<int64_t>(id) - 1 + min : 0;
      if (type == VB2_INTEGER) reinterpret_cast<int32_t*>(values)[i] = static_cast<int32_t>(v);
      else if (type == VB2_BOOLEAN) reinterpret_cast<uint8_t*>(values)[i] = static_cast<uint8_t>(v);
      else reinterpret_cast<int64_t*>(values)[i] = v;
    }
    const unsigned word = __ballot_sync(0xffffffffu, valid);
    if (lane == 0) valid_words[w] = word;
  }
}

__global__ void minmax_kernel(const __grid_constant__ vb2_column c, int64_t rows, int64_t* __restrict__ out3) {
  int64_t lo = INT64_MAX, hi = INT64_MIN, cnt = 0;
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < rows; r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t base;
    if (decode_row2(c, r, base)) continue;
    const int64_t v = key_value(c, base);
    lo = v < lo ? v : lo;
    hi = v > hi ? v : hi;
    ++cnt;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(reinterpret_cast<long long*>(out3), static_cast<long long>(lo));
    atomicMax(reinterpret_cast<long long*>(out3 + 1), static_cast<long long>(hi));
    atomicAdd(reinterpret_cast<unsigned long long*>(out3 + 2), static_cast<unsigned long long>(cnt));
  }
}
__global__ void minmax_init_kernel(int64_t* out3) {
  out3[0] = INT64_MAX;
  out3[1] = INT64_MIN;
  out3[2] = 0;
}

// Find-or-insert. One thread per row; the CAS on the key word both claims and publishes the slot.
__global__ void group_probe_kernel(const uint64_t* __restrict__ row_keys, const uint64_t* __restrict__ row_valid, int64_t n,
                                   uint64_t* __restrict__ table, uint64_t mask, int32_t* __restrict__ group_ids,
                                   int64_t* __restrict__ num_groups, int32_t* __restrict__ error_flag) {
  int64_t fresh = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (row_valid && !bit_at(row_valid, i)) { group_ids[i] = -1; continue; }
    const uint64_t key = row_keys[i];
    if (table == nullptr) {  // array mode: the normalized key is the slot
      group_ids[i] = key <= mask ? static_cast<int32_t>(key) : -1;
      if (key > mask) atomicCAS(error_flag, 0, 100);
      continue;
    }
    uint64_t slot = twang_mix64(key) & mask;
    int32_t found = -1;
    for (uint64_t probes = 0; probes <= mask; ++probes) {
      uint64_t cur = table[slot];
      if (cur == VB2_EMPTY_KEY) {
        cur = atomicCAS(reinterpret_cast<unsigned long long*>(table + slot), VB2_EMPTY_KEY, static_cast<unsigned long long>(key));
        if (cur == VB2_EMPTY_KEY) { ++fresh; found = static_cast<int32_t>(slot); break; }
      }
      if (cur == key) { found = static_cast<int32_t>(slot); break; }
      slot = (slot + 1) & mask;
    }
    if (found < 0) atomicCAS(error_flag, 0, 100);  // table full: the host sized it wrongly
    group_ids[i] = found;
  }
  fresh = warp_sum(fresh);
  if ((threadIdx.x & 31) == 0 && fresh) atomicAdd(reinterpret_cast<unsigned long long*>(num_groups), static_cast<unsigned long long>(fresh));
}

// ---- accumulator updates ----------------------------------------------------------------------
constexpr int kMaxAggs = 16;
struct AggArgs {
  vb2_agg_update a[kMaxAggs];
  int n;
};

__device__ __forceinline__ double input_as_f64(const vb2_agg_update& u, int64_t i) {
  switch (u.input_type) {
    case VB2_DOUBLE: return reinterpret_cast<const double*>(u.input)[i];
    case VB2_BIGINT: return static_cast<double>(reinterpret_cast<const int64_t*>(u.input)[i]);
    default: return static_cast<double>(reinterpret_cast<const int32_t*>(u.input)[i]);
  }
}
__device__ __forceinline__ int64_t input_as_i64(const vb2_agg_update& u, int64_t i) {
  switch (u.input_type) {
    case VB2_BIGINT: return reinterpret_cast<const int64_t*>(u.input)[i];
    case VB2_INTEGER: return reinterpret_cast<const int32_t*>(u.input)[i];
    default: return reinterpret_cast<const uint8_t*>(u.input)[i];
  }
}

__device__ __forceinline__ void atomic_min_f64(double* addr, double v, bool is_min) {
  unsigned long long* p = reinterpret_cast<unsigned long long*>(addr);
  unsigned long long old = *p;
  for (;;) {
    const double cur = __longlong_as_double(static_cast<long long>(old));
    const bool better = is_min ? lt_f64(v, cur) : gt_f64(v, cur);
    if (!better) return;
    const unsigned long long seen = atomicCAS(p, old, static_cast<unsigned long long>(__double_as_longlong(v)));
    if (seen == old) return;
    old = seen;
  }
}

// General path: one atomic per (row, aggregate). Sector-random for high-cardinality GROUP BY,
// which is the access pattern that bounds config 5.
__global__ void agg_update_atomic_kernel(const int32_t* __restrict__ group_ids, int64_t n, const __grid_constant__ AggArgs args,
                                         int32_t* __restrict__ error_flag) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t g = group_ids ? group_ids[i] : 0;
    if (g < 0) continue;
    for (int k = 0; k < args.n; ++k) {
      const vb2_agg_update& u = args.a[k];
      if (u.mask && !bit_at(u.mask, i)) continue;
      if (u.nulls && !bit_at(u.nulls, i)) continue;
      switch (u.kind) {
        case VB2_AGG_SUM_F64: atomicAdd(reinterpret_cast<double*>(u.acc) + g, input_as_f64(u, i)); break;
        case VB2_AGG_SUM_I64: case VB2_AGG_COUNT_MERGE: {
          const int64_t v = input_as_i64(u, i);
          const int64_t old = static_cast<int64_t>(atomicAdd(reinterpret_cast<unsigned long long*>(u.acc) + g, static_cast<unsigned long long>(v)));
          int64_t r;
          if (add_overflow_i64(old, v, &r)) atomicCAS(error_flag, 0, 1);
          break;
        }
        case VB2_AGG_COUNT: atomicAdd(reinterpret_cast<unsigned long long*>(u.acc) + g, 1ull); break;
        case VB2_AGG_MIN_F64: atomic_min_f64(reinterpret_cast<double*>(u.acc) + g, input_as_f64(u, i), true); break;
        case VB2_AGG_MAX_F64: atomic_min_f64(reinterpret_cast<double*>(u.acc) + g, input_as_f64(u, i), false); break;
        case VB2_AGG_MIN_I64: atomicMin(reinterpret_cast<long long*>(u.acc) + g, static_cast<long long>(input_as_i64(u, i))); break;
        case VB2_AGG_MAX_I64: atomicMax(reinterpret_cast<long long*>(u.acc) + g, static_cast<long long>(input_as_i64(u, i))); break;
        default: break;
      }
      if (u.nonnull && u.kind != VB2_AGG_COUNT) atomicAdd(reinterpret_cast<unsigned long long*>(u.nonnull) + g, 1ull);
    }
  }
}

// Tiny group-id spaces (<= 8 groups): same-address atomics would serialise in L2, so every thread
// keeps the groups in registers (predicated adds) and the block issues one atomic per group.
constexpr int kTinyG = 8;
template <int kKind>
__global__ void agg_update_tiny_kernel(const int32_t* __restrict__ group_ids, int64_t n, const __grid_constant__ vb2_agg_update u,
                                       int32_t* __restrict__ error_flag) {
  double fs[kTinyG];
  int64_t is[kTinyG], cnt[kTinyG];
#pragma unroll
  for (int g = 0; g < kTinyG; ++g) { fs[g] = 0.0; is[g] = 0; cnt[g] = 0; }
  bool ovf = false;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int32_t g = group_ids ? group_ids[i] : 0;
    if (u.mask && !bit_at(u.mask, i)) g = -1;
    if (u.nulls && !bit_at(u.nulls, i)) g = -1;
    if (g < 0) continue;
    double fv = 0.0;
    int64_t iv = 0;
    if (kKind == VB2_AGG_SUM_F64) fv = input_as_f64(u, i);
    if (kKind == VB2_AGG_SUM_I64 || kKind == VB2_AGG_COUNT_MERGE) iv = input_as_i64(u, i);
#pragma unroll
    for (int k = 0; k < kTinyG; ++k) {
      if (g == k) {
        cnt[k] += 1;
        if (kKind == VB2_AGG_SUM_F64) fs[k] = __dadd_rn(fs[k], fv);
        if (kKind == VB2_AGG_SUM_I64 || kKind == VB2_AGG_COUNT_MERGE) ovf |= add_overflow_i64(is[k], iv, &is[k]);
      }
    }
  }
  __shared__ double sf[256 / kWarp][kTinyG];
  __shared__ int64_t si[256 / kWarp][kTinyG];
  __shared__ int64_t sc[256 / kWarp][kTinyG];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < kTinyG; ++k) {
    const double f = warp_sum(fs[k]);
    // integer partials: detect overflow while combining
    int64_t v = is[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int64_t other = __shfl_xor_sync(0xffffffffu, v, o);
      ovf |= add_overflow_i64(v, other, &v);
    }
    const int64_t c = warp_sum(cnt[k]);
    if (lane == 0) { sf[warp][k] = f; si[warp][k] = v; sc[warp][k] = c; }
  }
  if (__any_sync(0xffffffffu, ovf) && lane == 0) atomicCAS(error_flag, 0, 1);
  __syncthreads();
  if (threadIdx.x < kTinyG) {
    const int k = threadIdx.x;
    double f = 0.0;
    int64_t v = 0, c = 0;
    bool o2 = false;
    for (int w = 0; w < 256 / kWarp; ++w) {
      f = __dadd_rn(f, sf[w][k]);
      o2 |= add_overflow_i64(v, si[w][k], &v);
      c += sc[w][k];
    }
    if (c) {
      if (kKind == VB2_AGG_SUM_F64) atomicAdd(reinterpret_cast<double*>(u.acc) + k, f);
      if (kKind == VB2_AGG_SUM_I64 || kKind == VB2_AGG_COUNT_MERGE) {
        const int64_t old = static_cast<int64_t>(atomicAdd(reinterpret_cast<unsigned long long*>(u.acc) + k, static_cast<unsigned long long>(v)));
        int64_t r;
        o2 |= add_overflow_i64(old, v, &r);
      }
      if (kKind == VB2_AGG_COUNT) atomicAdd(reinterpret_cast<unsigned long long*>(u.acc) + k, static_cast<unsigned long long>(c));
      if (u.nonnull && kKind != VB2_AGG_COUNT) atomicAdd(reinterpret_cast<unsigned long long*>(u.nonnull) + k, static_cast<unsigned long long>(c));
    }
    if (o2) atomicCAS(error_flag, 0, 1);
  }
}

// Re-encodes the keys of occupied slots for a new layout (value ranges grew): decodes the
// per-column ids with the old (mult, range), re-bases them on the new mins and packs again.
struct RekeyArgs {
  int n;
  int64_t old_min[kMaxNormCols], new_min[kMaxNormCols];
  uint64_t old_mult[kMaxNormCols], old_range[kMaxNormCols], new_mult[kMaxNormCols];
  int old_null_reserved[kMaxNormCols];
};
__global__ void rekey_kernel(const uint64_t* __restrict__ table_keys, const int32_t* __restrict__ slots, int64_t n,
                             const __grid_constant__ RekeyArgs a, uint64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint64_t key = table_keys ? table_keys[slots[i]] : static_cast<uint64_t>(slots[i]);
    uint64_t nk = 0;
    for (int k = 0; k < a.n; ++k) {
      const uint64_t id = (key / a.old_mult[k]) % a.old_range[k];
      const bool is_null = a.old_null_reserved[k] && id == 0;
      const uint64_t nid = is_null ? 0 : static_cast<uint64_t>(static_cast<int64_t>(id) - 1 + a.old_min[k] - a.new_min[k] + 1);
      nk += nid * a.new_mult[k];
    }
    out[i] = nk;
  }
}

// ---- occupied slots ---------------------------------------------------------------------------
__global__ void occupied_bits_kernel(const uint64_t* __restrict__ table, int64_t capacity, uint32_t* __restrict__ bits) {
  const int64_t nwords = (capacity + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t s = (w << 5) + lane;
    const bool occ = s < capacity && table[s] != VB2_EMPTY_KEY;
    const unsigned word = __ballot_sync(0xffffffffu, occ);
    if (lane == 0) bits[w] = word;
  }
}

static unsigned grid_for(int64_t n, int threads, int per_sm = 8) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(device_sm_count()) * per_sm;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace vb2

using namespace vb2;

extern "C" {

int vb2k_normalize_keys(const vb2_column* cols, int32_t ncols, const int64_t* mins, const uint64_t* mults, const uint64_t* ranges,
                        int32_t nulls_invalid, const int32_t* sel, int64_t n, uint64_t* keys_out, uint64_t* valid_out, void* stream) {
  if (ncols < 1 || ncols > kMaxNormCols) return fail_msg(VB2_ERR_UNSUPPORTED, "normalize_keys: 1..4 key columns");
  if (n <= 0) return VB2_OK;
  NormArgs a;
  a.n = ncols;
  a.nulls_invalid = nulls_invalid;
  a.check_ranges = ranges != nullptr;
  for (int i = 0; i < ncols; ++i) {
    if (cols[i].type == VB2_DOUBLE || cols[i].type == VB2_VARCHAR) return fail_msg(VB2_ERR_INVALID, "normalize_keys: integer-typed key columns expected");
    a.c[i] = cols[i];
    a.mins[i] = mins[i];
    a.mults[i] = mults[i];
    a.ranges[i] = ranges ? ranges[i] : 0;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (valid_out) VB2_CUDA_OK(cudaMemsetAsync(valid_out + ((n + 63) >> 6) - 1, 0, 8, st));
  normalize_keys_kernel<<<grid_for(n, 256), 256, 0, st>>>(a, sel, n, keys_out, reinterpret_cast<uint32_t*>(valid_out));
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_denormalize_keys(const uint64_t* table_keys, const int32_t* slots, int64_t n, int64_t min, uint64_t mult,
                          uint64_t range, int32_t null_reserved, int32_t type, void* values, uint64_t* valid, void* stream) {
  if (n <= 0) return VB2_OK;
  denormalize_keys_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      table_keys, slots, n, min, mult, range, null_reserved, type, values, reinterpret_cast<uint32_t*>(valid));
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_column_minmax(const vb2_column* col, int64_t rows, int64_t* out3, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  minmax_init_kernel<<<1, 1, 0, st>>>(out3);
  if (rows > 0) minmax_kernel<<<grid_for(rows, 256), 256, 0, st>>>(*col, rows, out3);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_probe(const uint64_t* row_keys, const uint64_t* row_valid, int64_t n, uint64_t* table_keys,
                     int64_t capacity, int32_t* group_ids, int64_t* num_groups, int32_t* error_flag, void* stream) {
  if (capacity <= 0 || (table_keys && (capacity & (capacity - 1)))) return fail_msg(VB2_ERR_INVALID, "group_probe: capacity must be a power of two");
  if (capacity > (1ll << 31)) return fail_msg(VB2_ERR_UNSUPPORTED, "group_probe: capacity above 2^31 slots");
  if (n <= 0) return VB2_OK;
  group_probe_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      row_keys, row_valid, n, table_keys, static_cast<uint64_t>(capacity - 1), group_ids, num_groups, error_flag);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_agg_update(const int32_t* group_ids, int64_t n, int64_t capacity, const vb2_agg_update* aggs, int32_t naggs,
                    int32_t* error_flag, void* stream) {
  if (naggs < 0 || naggs > kMaxAggs) return fail_msg(VB2_ERR_UNSUPPORTED, "agg_update: at most 16 aggregates per call");
  if (n <= 0 || naggs == 0) return VB2_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AggArgs rest;
  rest.n = 0;
  for (int i = 0; i < naggs; ++i) {
    const vb2_agg_update& u = aggs[i];
    const bool tiny = capacity <= kTinyG && (u.kind == VB2_AGG_SUM_F64 || u.kind == VB2_AGG_SUM_I64 ||
                                             u.kind == VB2_AGG_COUNT || u.kind == VB2_AGG_COUNT_MERGE);
    if (!tiny) { rest.a[rest.n++] = u; continue; }
    const unsigned grid = grid_for(n, 256, 4);
    switch (u.kind) {
      case VB2_AGG_SUM_F64: agg_update_tiny_kernel<VB2_AGG_SUM_F64><<<grid, 256, 0, st>>>(group_ids, n, u, error_flag); break;
      case VB2_AGG_SUM_I64: agg_update_tiny_kernel<VB2_AGG_SUM_I64><<<grid, 256, 0, st>>>(group_ids, n, u, error_flag); break;
      case VB2_AGG_COUNT_MERGE: agg_update_tiny_kernel<VB2_AGG_COUNT_MERGE><<<grid, 256, 0, st>>>(group_ids, n, u, error_flag); break;
      default: agg_update_tiny_kernel<VB2_AGG_COUNT><<<grid, 256, 0, st>>>(group_ids, n, u, error_flag); break;
    }
  }
  if (rest.n) agg_update_atomic_kernel<<<grid_for(n, 256), 256, 0, st>>>(group_ids, n, rest, error_flag);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_table_occupied(const uint64_t* table_keys, int64_t capacity, int32_t* slot_list, int64_t* count,
                        void* workspace, size_t workspace_bytes, void* stream) {
  // workspace = occupancy bitmap (capacity bits, 8-byte aligned) followed by the compaction scratch
  const size_t bitmap_bytes = static_cast<size_t>((capacity + 63) >> 6) * 8;
  const size_t need = bitmap_bytes + vb2k_bits_to_indices_workspace(capacity);
  if (workspace_bytes < need) return fail_msg(VB2_ERR_INVALID, "table_occupied: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint64_t* bits = reinterpret_cast<uint64_t*>(workspace);
  VB2_CUDA_OK(cudaMemsetAsync(bits, 0, bitmap_bytes, st));
  occupied_bits_kernel<<<grid_for(capacity, 256), 256, 0, st>>>(table_keys, capacity, reinterpret_cast<uint32_t*>(bits));
  VB2_CUDA_OK(cudaGetLastError());
  return vb2k_bits_to_indices(bits, capacity, slot_list, count, reinterpret_cast<char*>(workspace) + bitmap_bytes,
                              workspace_bytes - bitmap_bytes, stream);
}

int vb2k_rekey(const uint64_t* table_keys, const int32_t* slots, int64_t n, int32_t ncols, const int64_t* old_mins,
               const uint64_t* old_mults, const uint64_t* old_ranges, const int32_t* old_null_reserved, const int64_t* new_mins,
               const uint64_t* new_mults, uint64_t* keys_out, void* stream) {
  if (ncols < 1 || ncols > kMaxNormCols) return fail_msg(VB2_ERR_UNSUPPORTED, "rekey: 1..4 key columns");
  if (n <= 0) return VB2_OK;
  RekeyArgs a;
  a.n = ncols;
  for (int i = 0; i < ncols; ++i) {
    a.old_min[i] = old_mins[i]; a.new_min[i] = new_mins[i];
    a.old_mult[i] = old_mults[i]; a.old_range[i] = old_ranges[i]; a.new_mult[i] = new_mults[i];
    a.old_null_reserved[i] = old_null_reserved ? old_null_reserved[i] : 1;
  }
  rekey_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(table_keys, slots, n, a, keys_out);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

size_t vb2k_table_occupied_workspace(int64_t capacity) {
  return static_cast<size_t>((capacity + 63) >> 6) * 8 + vb2k_bits_to_indices_workspace(capacity);
}

}  // extern "C"
