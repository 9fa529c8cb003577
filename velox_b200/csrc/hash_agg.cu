// Generic aggregation kernels: normalized keys, find-or-insert group table, accumulator updates.
//
// B200-native take on exec::HashTable + RowContainer for GROUP BY:
//   * keys of a row are packed into ONE 64-bit normalized key from per-column value ids
//     (id = v - min + 1, 0 = NULL; exec/VectorHasher.h:523-585) — the reference's kNormalizedKey
//     mode (exec/HashTable.cpp:523) is the only mode needed because ranges come from a device
//     min/max pass over the whole batch, not from 1K-row increments;
//   * the table is row-wise like RowContainer (exec/RowContainer.h): [key | accumulators...] in
//     whole 32-byte sectors, open addressing + linear probing on twang_mix64(key); find-or-insert
//     and every accumulator update of an input row happen in ONE kernel and touch one row;
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
// Flat NULL-free BIGINT column (the common shape of a high-cardinality key): 128-bit loads, eight values in flight per thread.
__global__ void __launch_bounds__(256) minmax_flat_i64_kernel(const int64_t* __restrict__ v, int64_t rows, int64_t* __restrict__ out3) {
  int64_t lo = INT64_MAX, hi = INT64_MIN;
  const int64_t pairs = rows >> 1;
  const longlong2* v2 = reinterpret_cast<const longlong2*>(v);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (; i + 3 * stride < pairs; i += 4 * stride) {
    longlong2 a[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) a[u] = __ldcs(v2 + i + u * stride);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      lo = a[u].x < lo ? a[u].x : lo;
      hi = a[u].x > hi ? a[u].x : hi;
      lo = a[u].y < lo ? a[u].y : lo;
      hi = a[u].y > hi ? a[u].y : hi;
    }
  }
  for (; i < pairs; i += stride) {
    const longlong2 a = v2[i];
    lo = a.x < lo ? a.x : lo;
    hi = a.x > hi ? a.x : hi;
    lo = a.y < lo ? a.y : lo;
    hi = a.y > hi ? a.y : hi;
  }
  if ((rows & 1) && blockIdx.x == 0 && threadIdx.x == 0) {
    const int64_t x = v[rows - 1];
    lo = x < lo ? x : lo;
    hi = x > hi ? x : hi;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const int64_t l2 = __shfl_xor_sync(0xffffffffu, lo, o), h2 = __shfl_xor_sync(0xffffffffu, hi, o);
    lo = l2 < lo ? l2 : lo;
    hi = h2 > hi ? h2 : hi;
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMin(reinterpret_cast<long long*>(out3), static_cast<long long>(lo));
    atomicMax(reinterpret_cast<long long*>(out3 + 1), static_cast<long long>(hi));
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) out3[2] = rows;  // no NULLs: every row counts
}
__global__ void minmax_init_kernel(int64_t* out3) {
  out3[0] = INT64_MAX;
  out3[1] = INT64_MIN;
  out3[2] = 0;
}

// Dictionary codes of a DICTIONARY / CONSTANT column as a dense int32 column plus validity bytes
// (NULL wrapper rows and NULL dictionary entries -> 0): the form VARCHAR columns take in an exchange.
__global__ void dictionary_codes_kernel(const __grid_constant__ vb2_column c, int64_t n, int32_t* __restrict__ codes, uint8_t* __restrict__ valid) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t base;
    const bool is_null = decode_row2(c, i, base);
    codes[i] = is_null ? 0 : static_cast<int32_t>(base);
    if (valid) valid[i] = is_null ? 0 : 1;
  }
}

// Array-mode join build straight from the key column (normalize + insert in one pass): slot =
// v - lo + 1, the value id vb2k_normalize_keys would produce for a single key with min = lo.
__global__ void join_build_array_direct_kernel(int32_t* __restrict__ head, int32_t* __restrict__ next, int64_t capacity,
                                               const __grid_constant__ vb2_column c, int64_t lo, int64_t n, int32_t* __restrict__ flags) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < n; r += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int64_t base;
    if (decode_row2(c, r, base)) continue;  // NULL keys are not inserted (exec/HashBuild.cpp:475-479)
    const int64_t slot = key_value(c, base) - lo + 1;
    if (slot < 1 || slot >= capacity) { atomicCAS(flags, 0, 101); continue; }
    const int32_t prev = atomicExch(head + slot, static_cast<int32_t>(r + 1));
    next[r] = prev;
    if (prev != 0) atomicCAS(flags + 1, 0, 1);  // duplicate build keys present (informational)
  }
}

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

// ---- group table ---------------------------------------------------------------------------------
// Row-wise group storage (the role of exec::RowContainer under exec::HashTable for GROUP BY):
// `capacity` rows of `row_words` 8-byte words. Word 0 is the occupancy word — the normalized key in
// hash mode (VB2_EMPTY_KEY = free slot), "rows seen" in array / global mode (0 = free) — the
// remaining words are accumulators and non-null counters. A row is a whole number of 32-byte
// sectors whenever it has more than two words, so one input row touches ONE random sector group
// (key compare + every accumulator) instead of one per accumulator array.
struct TableView {
  uint64_t* rows;
  uint64_t mask;      // hash mode: capacity - 1; array mode: capacity - 1 is the largest valid key
  int32_t w;
  int32_t hash;
  int32_t shift;      // hash mode: home slot = twang_mix64(key) >> shift — the TOP bits of the hash, so that rows
                      // ordered by their hash's top bits (radix_partition.cu) walk the table slice by slice
};
__device__ __forceinline__ int32_t top_bits_shift(int64_t capacity) { return capacity > 1 ? 64 - (63 - __clzll(capacity)) : 63; }
__device__ __forceinline__ TableView view_of(const vb2_group_table& t) {
  return TableView{t.rows, static_cast<uint64_t>(t.capacity - 1), t.row_words, t.hash_mode, top_bits_shift(t.capacity)};
}

// Slot of `key` (inserted if absent), -1 if the table is full / the key is outside the array.
__device__ __forceinline__ int64_t find_or_insert(const TableView& t, uint64_t key, int64_t& fresh) {
  if (!t.hash) return key <= t.mask ? static_cast<int64_t>(key) : -1;
  uint64_t slot = (twang_mix64(key) >> t.shift) & t.mask;
  for (uint64_t probes = 0; probes <= t.mask; ++probes) {
    uint64_t* p = t.rows + slot * t.w;
    uint64_t cur = *reinterpret_cast<volatile uint64_t*>(p);
    if (cur == VB2_EMPTY_KEY) {
      cur = atomicCAS(reinterpret_cast<unsigned long long*>(p), VB2_EMPTY_KEY, static_cast<unsigned long long>(key));
      if (cur == VB2_EMPTY_KEY) { ++fresh; return static_cast<int64_t>(slot); }
    }
    if (cur == key) return static_cast<int64_t>(slot);
    slot = (slot + 1) & t.mask;
  }
  return -1;
}

constexpr int kMaxAggs = 16;
constexpr int64_t kSerialRows = 256;     // batches this small are accumulated in input order by one warp
constexpr int64_t kAtomicRows = 1 << 16;  // up to here contention on a tiny table costs less than one pass per aggregate
struct AggArgs {
  vb2_agg_update a[kMaxAggs];
  int n;
};

// Input position of row i (identity or through the dictionary wrap), -1 if the row does not
// contribute (masked out, NULL wrapper row or NULL wrapped value).
__device__ __forceinline__ int64_t input_pos(const vb2_agg_update& u, int64_t i) {
  if (u.mask && !bit_at(u.mask, i)) return -1;
  if (u.nulls && !bit_at(u.nulls, i)) return -1;
  const int64_t j = u.indices ? u.indices[i] : i;
  if (u.base_nulls && !bit_at(u.base_nulls, j)) return -1;
  return j;
}

__device__ __forceinline__ void apply_update(const vb2_agg_update& u, int64_t row_index, uint64_t* row, int32_t* error_flag) {
  const int64_t i = input_pos(u, row_index);
  if (i < 0) return;
  uint64_t* acc = row + u.acc_word;
  switch (u.kind) {
    case VB2_AGG_SUM_F64: atomicAdd(reinterpret_cast<double*>(acc), input_as_f64(u, i)); break;
    case VB2_AGG_SUM_I64: case VB2_AGG_COUNT_MERGE: {
      const int64_t v = input_as_i64(u, i);
      const int64_t old = static_cast<int64_t>(atomicAdd(reinterpret_cast<unsigned long long*>(acc), static_cast<unsigned long long>(v)));
      int64_t r;
      if (add_overflow_i64(old, v, &r)) atomicCAS(error_flag, 0, 1);
      break;
    }
    case VB2_AGG_COUNT: atomicAdd(reinterpret_cast<unsigned long long*>(acc), 1ull); break;
    case VB2_AGG_MIN_F64: atomic_min_f64(reinterpret_cast<double*>(acc), input_as_f64(u, i), true); break;
    case VB2_AGG_MAX_F64: atomic_min_f64(reinterpret_cast<double*>(acc), input_as_f64(u, i), false); break;
    case VB2_AGG_MIN_I64: atomicMin(reinterpret_cast<long long*>(acc), static_cast<long long>(input_as_i64(u, i))); break;
    case VB2_AGG_MAX_I64: atomicMax(reinterpret_cast<long long*>(acc), static_cast<long long>(input_as_i64(u, i))); break;
    default: break;
  }
  if (u.nonnull_word >= 0 && u.kind != VB2_AGG_COUNT) atomicAdd(reinterpret_cast<unsigned long long*>(row + u.nonnull_word), 1ull);
}

// General path, ONE kernel per batch: find-or-insert the row's group, then every accumulator update
// lands in that group's row (same sectors as the key that was just compared).
__global__ void group_update_kernel(const __grid_constant__ vb2_group_table tab, const uint64_t* __restrict__ row_keys,
                                    const uint64_t* __restrict__ row_valid, int64_t n, const __grid_constant__ AggArgs args,
                                    int64_t* __restrict__ num_groups, int32_t* __restrict__ error_flag) {
  const TableView t = view_of(tab);
  int64_t fresh = 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // (An L2 prefetch of the home row of inputs 1-3 iterations ahead was measured and rejected:
  // 149 ms vs 119 ms per 1 B rows at 100 M groups — the kernel is bound by random-sector DRAM
  // throughput incl. page-walk traffic, not by the latency of one dependent access per thread.)
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    if (row_valid && !bit_at(row_valid, i)) continue;
    const int64_t slot = row_keys ? find_or_insert(t, row_keys[i], fresh) : 0;
    if (slot < 0) { atomicCAS(error_flag, 0, 100); continue; }  // the host sized the table wrongly
    uint64_t* row = t.rows + slot * t.w;
    if (!t.hash && *reinterpret_cast<volatile uint64_t*>(row) == 0) *reinterpret_cast<volatile uint64_t*>(row) = 1;  // occupancy mark (idempotent)
    for (int k = 0; k < args.n; ++k) apply_update(args.a[k], i, row, error_flag);
  }
  fresh = warp_sum(fresh);
  if ((threadIdx.x & 31) == 0 && fresh && num_groups) atomicAdd(reinterpret_cast<unsigned long long*>(num_groups), static_cast<unsigned long long>(fresh));
}

// Array-mode tables from a handful to a few thousand groups (GROUP BY on a low-cardinality column with
// NULLs / dictionary inputs — shapes the fused pipelines do not take): per-row atomics on a global
// table of 50 rows serialise in L2 on 50 addresses for the whole GPU. Every block accumulates into a
// private copy of the table in shared memory instead (same update code: the atomics land in the SM's
// own banks) and merges its copy into the global table once, with one atomic per touched word.
constexpr int kSmemTableBytes = 40 * 1024;
__device__ __forceinline__ uint64_t identity_of(int kind) {
  switch (kind) {
    case VB2_AGG_MIN_F64: return 0x7ff8000000000000ull;  // NaN: the largest value
    case VB2_AGG_MAX_F64: return 0xfff0000000000000ull;  // -inf
    case VB2_AGG_MIN_I64: return static_cast<uint64_t>(INT64_MAX);
    case VB2_AGG_MAX_I64: return static_cast<uint64_t>(INT64_MIN);
    default: return 0;
  }
}
__global__ void __launch_bounds__(256) group_update_smem_kernel(const __grid_constant__ vb2_group_table tab, const uint64_t* __restrict__ row_keys,
                                                                const uint64_t* __restrict__ row_valid, int64_t n,
                                                                const __grid_constant__ AggArgs args, int32_t* __restrict__ error_flag) {
  extern __shared__ __align__(16) uint64_t srows[];  // [capacity][row_words]
  const int w = tab.row_words;
  const int cap = static_cast<int>(tab.capacity);
  for (int i = threadIdx.x; i < cap * w; i += blockDim.x) srows[i] = 0;
  __syncthreads();
  for (int k = 0; k < args.n; ++k) {
    const uint64_t id = identity_of(args.a[k].kind);
    if (id)
      for (int slot = threadIdx.x; slot < cap; slot += blockDim.x) srows[slot * w + args.a[k].acc_word] = id;
  }
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    if (row_valid && !bit_at(row_valid, i)) continue;
    const uint64_t key = row_keys[i];
    if (key >= static_cast<uint64_t>(cap)) { atomicCAS(error_flag, 0, 100); continue; }
    uint64_t* row = srows + static_cast<int>(key) * w;
    if (*reinterpret_cast<volatile uint64_t*>(row) == 0) *reinterpret_cast<volatile uint64_t*>(row) = 1;
    for (int k = 0; k < args.n; ++k) apply_update(args.a[k], i, row, error_flag);
  }
  __syncthreads();
  for (int slot = threadIdx.x; slot < cap; slot += blockDim.x) {
    const uint64_t* mine = srows + slot * w;
    if (mine[0] == 0) continue;  // this block saw no row of the group
    uint64_t* row = tab.rows + static_cast<int64_t>(slot) * w;
    if (*reinterpret_cast<volatile uint64_t*>(row) == 0) *reinterpret_cast<volatile uint64_t*>(row) = 1;
    for (int k = 0; k < args.n; ++k) {
      const vb2_agg_update& u = args.a[k];
      const uint64_t v = mine[u.acc_word];
      uint64_t* acc = row + u.acc_word;
      switch (u.kind) {
        case VB2_AGG_SUM_F64: atomicAdd(reinterpret_cast<double*>(acc), __longlong_as_double(static_cast<long long>(v))); break;
        case VB2_AGG_SUM_I64: case VB2_AGG_COUNT_MERGE: {
          const int64_t x = static_cast<int64_t>(v);
          const int64_t old = static_cast<int64_t>(atomicAdd(reinterpret_cast<unsigned long long*>(acc), static_cast<unsigned long long>(x)));
          int64_t r;
          if (add_overflow_i64(old, x, &r)) atomicCAS(error_flag, 0, 1);
          break;
        }
        case VB2_AGG_COUNT: atomicAdd(reinterpret_cast<unsigned long long*>(acc), static_cast<unsigned long long>(v)); break;
        case VB2_AGG_MIN_F64: atomic_min_f64(reinterpret_cast<double*>(acc), __longlong_as_double(static_cast<long long>(v)), true); break;
        case VB2_AGG_MAX_F64: atomic_min_f64(reinterpret_cast<double*>(acc), __longlong_as_double(static_cast<long long>(v)), false); break;
        case VB2_AGG_MIN_I64: atomicMin(reinterpret_cast<long long*>(acc), static_cast<long long>(v)); break;
        case VB2_AGG_MAX_I64: atomicMax(reinterpret_cast<long long*>(acc), static_cast<long long>(v)); break;
        default: break;
      }
      if (u.nonnull_word >= 0 && u.kind != VB2_AGG_COUNT && mine[u.nonnull_word])
        atomicAdd(reinterpret_cast<unsigned long long*>(row + u.nonnull_word), static_cast<unsigned long long>(mine[u.nonnull_word]));
    }
  }
}

// ---- keyed hash mode (kHash) ----------------------------------------------------------------------
struct KeyedCols {
  vb2_column c[VB2_KEYED_MAX_KEYS];
  int n;
};
__device__ __forceinline__ uint64_t canonical_f64_bits(double v) {
  if (isnan(v)) return 0x7ff8000000000000ull;
  if (v == 0.0) return 0;  // -0 and +0 are one key
  return static_cast<uint64_t>(__double_as_longlong(v));
}
__device__ __forceinline__ uint64_t key_word_of(const vb2_column& c, int64_t base) {
  if (c.type == VB2_DOUBLE) return canonical_f64_bits(reinterpret_cast<const double*>(c.values)[base]);
  return static_cast<uint64_t>(key_value(c, base));
}
__device__ __forceinline__ uint64_t keyed_hash(const uint64_t* kw, uint64_t nullmask, int nkeys) {
  uint64_t h = 0;
  for (int k = 0; k < nkeys; ++k) {
    const uint64_t hk = (nullmask >> k) & 1u ? kNullHash : twang_mix64(kw[k]);
    h = k == 0 ? hk : hash_mix(h, hk);
  }
  return h >> 1;  // 63 bits: the state word keeps a ready bit
}
// Slot of the key (inserted if absent) in a keyed table, -1 when the table is full.
__device__ __forceinline__ int64_t find_or_insert_keyed(uint64_t* rows, uint64_t mask, int w, int nkeys, const uint64_t* kw, uint64_t nullmask, int64_t& fresh) {
  const uint64_t h63 = keyed_hash(kw, nullmask, nkeys);
  uint64_t slot = twang_mix64(h63) & mask;
  for (uint64_t probes = 0; probes <= mask; ++probes) {
    uint64_t* row = rows + slot * w;
    uint64_t st = *reinterpret_cast<volatile uint64_t*>(row);
    if (st == VB2_EMPTY_KEY) {
      st = atomicCAS(reinterpret_cast<unsigned long long*>(row), VB2_EMPTY_KEY, static_cast<unsigned long long>(h63 << 1));
      if (st == VB2_EMPTY_KEY) {
        for (int k = 0; k < nkeys; ++k) row[1 + k] = kw[k];
        row[1 + nkeys] = nullmask;
        __threadfence();
        *reinterpret_cast<volatile uint64_t*>(row) = (h63 << 1) | 1u;  // publish
        ++fresh;
        return static_cast<int64_t>(slot);
      }
    }
    if ((st >> 1) == h63) {
      while (!(st & 1u)) st = *reinterpret_cast<volatile uint64_t*>(row);  // claimed, keys not published yet
      __threadfence();
      bool same = reinterpret_cast<volatile uint64_t*>(row)[1 + nkeys] == nullmask;
      for (int k = 0; k < nkeys && same; ++k) same = reinterpret_cast<volatile uint64_t*>(row)[1 + k] == kw[k];
      if (same) return static_cast<int64_t>(slot);
    }
    slot = (slot + 1) & mask;
  }
  return -1;
}
__global__ void group_update_keyed_kernel(const __grid_constant__ vb2_group_table tab, const __grid_constant__ KeyedCols keys, int64_t n,
                                          const __grid_constant__ AggArgs args, int64_t* __restrict__ num_groups, int32_t* __restrict__ error_flag) {
  const uint64_t mask = static_cast<uint64_t>(tab.capacity - 1);
  int64_t fresh = 0;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += stride) {
    uint64_t kw[VB2_KEYED_MAX_KEYS];
    uint64_t nullmask = 0;
    for (int k = 0; k < keys.n; ++k) {
      int64_t base;
      const bool is_null = decode_row2(keys.c[k], i, base);
      kw[k] = is_null ? 0 : key_word_of(keys.c[k], base);
      nullmask |= static_cast<uint64_t>(is_null) << k;
    }
    const int64_t slot = find_or_insert_keyed(tab.rows, mask, tab.row_words, keys.n, kw, nullmask, fresh);
    if (slot < 0) { atomicCAS(error_flag, 0, 100); continue; }
    uint64_t* row = tab.rows + slot * tab.row_words;
    for (int k = 0; k < args.n; ++k) apply_update(args.a[k], i, row, error_flag);
  }
  fresh = warp_sum(fresh);
  if ((threadIdx.x & 31) == 0 && fresh && num_groups) atomicAdd(reinterpret_cast<unsigned long long*>(num_groups), static_cast<unsigned long long>(fresh));
}
// Slot of a key tuple already in a keyed table whose rows are all published (no concurrent inserts), -1 when absent.
__device__ __forceinline__ int64_t find_keyed(const uint64_t* rows, uint64_t mask, int w, int nkeys, const uint64_t* kw) {
  const uint64_t h63 = keyed_hash(kw, 0, nkeys);
  uint64_t slot = twang_mix64(h63) & mask;
  for (uint64_t probes = 0; probes <= mask; ++probes) {
    const uint64_t* row = rows + slot * w;
    const uint64_t st = row[0];
    if (st == VB2_EMPTY_KEY) return -1;
    if ((st >> 1) == h63) {
      bool same = row[1 + nkeys] == 0;
      for (int k = 0; k < nkeys && same; ++k) same = row[1 + k] == kw[k];
      if (same) return static_cast<int64_t>(slot);
    }
    slot = (slot + 1) & mask;
  }
  return -1;
}
// Join keys that do not normalize into one word: the id of a key tuple is its slot in a keyed table. One warp
// per 32 consecutive rows, so the valid bits of those rows are one ballot word.
__global__ void keyed_key_ids_kernel(const __grid_constant__ vb2_group_table tab, const __grid_constant__ KeyedCols keys, int64_t n, int insert,
                                     uint64_t* __restrict__ ids, uint32_t* __restrict__ valid, int64_t* __restrict__ num_groups,
                                     int32_t* __restrict__ error_flag) {
  const uint64_t mask = static_cast<uint64_t>(tab.capacity - 1);
  const int64_t nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  int64_t fresh = 0;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t r = (w << 5) + lane;
    bool ok = false;
    uint64_t id = 0;
    if (r < n) {
      uint64_t kw[VB2_KEYED_MAX_KEYS];
      bool any_null = false;
      for (int k = 0; k < keys.n; ++k) {
        int64_t base;
        const bool is_null = decode_row2(keys.c[k], r, base);
        kw[k] = is_null ? 0 : key_word_of(keys.c[k], base);
        any_null |= is_null;
      }
      if (!any_null) {
        const int64_t slot = insert ? find_or_insert_keyed(tab.rows, mask, tab.row_words, keys.n, kw, 0, fresh)
                                    : find_keyed(tab.rows, mask, tab.row_words, keys.n, kw);
        if (slot >= 0) { ok = true; id = static_cast<uint64_t>(slot); }
        else if (insert) atomicCAS(error_flag, 0, 100);
      }
      ids[r] = id;
    }
    const unsigned word = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) valid[w] = word;
  }
  fresh = warp_sum(fresh);
  if (lane == 0 && fresh && num_groups) atomicAdd(reinterpret_cast<unsigned long long*>(num_groups), static_cast<unsigned long long>(fresh));
}
// Re-inserts groups into a (bigger) keyed table; keys are distinct, so the row belongs to this thread.
struct DecodeArgs {
  int n;
  int64_t mins[VB2_KEYED_MAX_KEYS];
  uint64_t mults[VB2_KEYED_MAX_KEYS], ranges[VB2_KEYED_MAX_KEYS];
  int null_reserved[VB2_KEYED_MAX_KEYS];
  int from_keyed;   // source rows already hold key words
  int word_shift;   // accumulator word w of the source -> w + word_shift
};
__global__ void group_move_keyed_kernel(const __grid_constant__ vb2_group_table from, const int32_t* __restrict__ slots, int64_t n,
                                        const __grid_constant__ DecodeArgs d, const __grid_constant__ vb2_group_table to, int64_t* __restrict__ num_groups,
                                        int32_t* __restrict__ error_flag) {
  const uint64_t mask = static_cast<uint64_t>(to.capacity - 1);
  int64_t fresh = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t s = slots[i];
    const uint64_t* src = from.rows + s * from.row_words;
    uint64_t kw[VB2_KEYED_MAX_KEYS];
    uint64_t nullmask = 0;
    int first_acc;
    if (d.from_keyed) {
      for (int k = 0; k < d.n; ++k) kw[k] = src[1 + k];
      nullmask = src[1 + d.n];
      first_acc = d.n + 2;
    } else {
      const uint64_t key = from.hash_mode ? src[0] : static_cast<uint64_t>(s);
      for (int k = 0; k < d.n; ++k) {
        const uint64_t id = (key / d.mults[k]) % d.ranges[k];
        const bool is_null = d.null_reserved[k] && id == 0;
        kw[k] = is_null ? 0 : static_cast<uint64_t>(static_cast<int64_t>(id) - 1 + d.mins[k]);
        nullmask |= static_cast<uint64_t>(is_null) << k;
      }
      first_acc = 1;
    }
    const int64_t slot = find_or_insert_keyed(to.rows, mask, to.row_words, d.n, kw, nullmask, fresh);
    if (slot < 0) { atomicCAS(error_flag, 0, 100); continue; }
    uint64_t* dst = to.rows + slot * to.row_words;
    for (int w = first_acc; w < from.row_words && w + d.word_shift < to.row_words; ++w) dst[w + d.word_shift] = src[w];
  }
  fresh = warp_sum(fresh);
  if ((threadIdx.x & 31) == 0 && fresh && num_groups) atomicAdd(reinterpret_cast<unsigned long long*>(num_groups), static_cast<unsigned long long>(fresh));
}

// Rows in radix-partition order (radix_partition.cu): partition p's rows only touch slice p of the
// table (the home slot is the top bits of the hash), so the grid walks the partitions in lock step —
// slice p + 1 is prefetched into L2 with full-line reads while the rows of partition p are folded into
// slice p, which is L2 resident by then. DRAM sees the table once per batch, sequentially, instead of
// one random 32-byte sector per input row. Cooperative launch (all blocks co-resident): the per-
// partition barrier is a monotonic counter.
__device__ __forceinline__ void grid_barrier(unsigned int* counter, unsigned int nblocks, unsigned int& generation) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    const unsigned int target = (generation + 1u) * nblocks;
    while (*reinterpret_cast<volatile unsigned int*>(counter) < target) {}
    __threadfence();
  }
  ++generation;
  __syncthreads();
}
__global__ void __launch_bounds__(256)
group_update_partitioned_kernel(const __grid_constant__ vb2_group_table tab, const uint64_t* __restrict__ row_keys, const int64_t* __restrict__ part_start,
                                int nparts, const __grid_constant__ AggArgs args, int64_t* __restrict__ num_groups, int32_t* __restrict__ error_flag,
                                unsigned int* __restrict__ barrier) {
  const TableView t = view_of(tab);
  const int64_t gtid = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  const int64_t gthreads = static_cast<int64_t>(gridDim.x) * blockDim.x;
  const int64_t slice_lines = (tab.capacity / nparts) * tab.row_words * 8 / 128;  // 128-byte lines per table slice
  const char* table_bytes = reinterpret_cast<const char*>(tab.rows);
  auto prefetch_slice = [&](int p) {
    const char* base = table_bytes + static_cast<int64_t>(p) * slice_lines * 128;
    for (int64_t l = gtid; l < slice_lines; l += gthreads) asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(base + l * 128));
  };
  unsigned int generation = 0;
  int64_t fresh = 0;
  prefetch_slice(0);
  grid_barrier(barrier, gridDim.x, generation);
  for (int p = 0; p < nparts; ++p) {
    if (p + 1 < nparts) prefetch_slice(p + 1);
    const int64_t r0 = part_start[p], r1 = part_start[p + 1];
    for (int64_t i = r0 + gtid; i < r1; i += gthreads) {
      const int64_t slot = find_or_insert(t, row_keys[i], fresh);
      if (slot < 0) { atomicCAS(error_flag, 0, 100); continue; }
      uint64_t* row = t.rows + slot * t.w;
      for (int k = 0; k < args.n; ++k) apply_update(args.a[k], i, row, error_flag);
    }
    grid_barrier(barrier, gridDim.x, generation);
  }
  fresh = warp_sum(fresh);
  if ((threadIdx.x & 31) == 0 && fresh && num_groups) atomicAdd(reinterpret_cast<unsigned long long*>(num_groups), static_cast<unsigned long long>(fresh));
}

// Very small batches (merging a handful of partial-aggregate rows, e.g. one row per group and GPU
// in front of a final aggregation): one warp walks the rows IN INPUT ORDER, lane k owning aggregate
// k, with plain read-modify-writes — the reference's sequential accumulation
// (functions/lib/aggregates/SimpleNumericAggregate.h:94-150), deterministic and free of atomics.
__device__ __forceinline__ void apply_update_plain(const vb2_agg_update& u, int64_t row_index, uint64_t* row, int32_t* error_flag) {
  const int64_t i = input_pos(u, row_index);
  if (i < 0) return;
  uint64_t* acc = row + u.acc_word;
  switch (u.kind) {
    case VB2_AGG_SUM_F64: *reinterpret_cast<double*>(acc) = __dadd_rn(*reinterpret_cast<double*>(acc), input_as_f64(u, i)); break;
    case VB2_AGG_SUM_I64: case VB2_AGG_COUNT_MERGE: {
      int64_t r;
      if (add_overflow_i64(static_cast<int64_t>(*acc), input_as_i64(u, i), &r)) atomicCAS(error_flag, 0, 1);
      *acc = static_cast<uint64_t>(r);
      break;
    }
    case VB2_AGG_COUNT: *acc += 1; break;
    case VB2_AGG_MIN_F64: { const double v = input_as_f64(u, i); if (lt_f64(v, *reinterpret_cast<double*>(acc))) *reinterpret_cast<double*>(acc) = v; break; }
    case VB2_AGG_MAX_F64: { const double v = input_as_f64(u, i); if (gt_f64(v, *reinterpret_cast<double*>(acc))) *reinterpret_cast<double*>(acc) = v; break; }
    case VB2_AGG_MIN_I64: { const int64_t v = input_as_i64(u, i); if (v < static_cast<int64_t>(*acc)) *acc = static_cast<uint64_t>(v); break; }
    case VB2_AGG_MAX_I64: { const int64_t v = input_as_i64(u, i); if (v > static_cast<int64_t>(*acc)) *acc = static_cast<uint64_t>(v); break; }
    default: break;
  }
  if (u.nonnull_word >= 0 && u.kind != VB2_AGG_COUNT) row[u.nonnull_word] += 1;
}
__global__ void __launch_bounds__(32)
group_update_serial_kernel(const __grid_constant__ vb2_group_table tab, const uint64_t* __restrict__ row_keys, const uint64_t* __restrict__ row_valid,
                           int64_t n, const __grid_constant__ AggArgs args, int64_t* __restrict__ num_groups, int32_t* __restrict__ error_flag) {
  const TableView t = view_of(tab);
  const int lane = threadIdx.x;
  int64_t fresh = 0;
  for (int64_t i = 0; i < n; ++i) {
    if (row_valid && !bit_at(row_valid, i)) continue;
    int64_t slot = 0;
    if (row_keys) {
      int64_t f = 0;
      if (lane == 0) slot = find_or_insert(t, row_keys[i], f);
      fresh += f;
      slot = __shfl_sync(0xffffffffu, slot, 0);
    }
    if (slot < 0) { if (lane == 0) atomicCAS(error_flag, 0, 100); continue; }
    uint64_t* row = t.rows + slot * t.w;
    if (!t.hash && lane == 0 && *row == 0) *row = 1;
    for (int k = lane; k < args.n; k += 32) apply_update_plain(args.a[k], i, row, error_flag);
    __syncwarp();
  }
  if (lane == 0 && fresh && num_groups) atomicAdd(reinterpret_cast<unsigned long long*>(num_groups), static_cast<unsigned long long>(fresh));
}

// Tiny array-mode tables (<= 8 rows): same-address atomics would serialise in L2, so every thread
// keeps the groups in registers (predicated adds) and the block issues one atomic per group.
// kKind == 0 is the occupancy pass (rows seen per group -> word 0).
constexpr int kTinyG = 8;
template <int kKind>
__global__ void group_update_tiny_kernel(const __grid_constant__ vb2_group_table tab, const uint64_t* __restrict__ row_keys,
                                         const uint64_t* __restrict__ row_valid, int64_t n, const __grid_constant__ vb2_agg_update u,
                                         int32_t* __restrict__ error_flag) {
  double fs[kTinyG];
  int64_t is[kTinyG], cnt[kTinyG];
#pragma unroll
  for (int g = 0; g < kTinyG; ++g) { fs[g] = 0.0; is[g] = 0; cnt[g] = 0; }
  bool ovf = false, bad = false;
  // Four independent rows per iteration, branch-free up to the accumulation: stage 1 issues the
  // key and index loads of all four rows, stage 2 the value loads they address, stage 3 adds —
  // 4 x 2 dependent loads in flight per thread instead of a 3-deep chain per row.
  constexpr int kU = 4;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i0 = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i0 < n; i0 += kU * stride) {
    int64_t ii[kU], jj[kU];
    uint64_t key[kU];
    bool ok[kU];
#pragma unroll
    for (int r = 0; r < kU; ++r) {
      const int64_t i = i0 + r * stride;
      ok[r] = i < n;
      ii[r] = ok[r] ? i : 0;
      key[r] = row_keys ? row_keys[ii[r]] : 0;
      jj[r] = ii[r];
      if (kKind != 0 && u.indices) {
        // indices of NULL wrapper rows are undefined: never follow them
        const bool wrapper_ok = !u.nulls || bit_at(u.nulls, ii[r]);
        jj[r] = wrapper_ok ? u.indices[ii[r]] : 0;
      }
    }
    double fv[kU];
    int64_t iv[kU];
#pragma unroll
    for (int r = 0; r < kU; ++r) {
      fv[r] = 0.0;
      iv[r] = 0;
      if (kKind == VB2_AGG_SUM_F64) fv[r] = input_as_f64(u, jj[r]);
      if (kKind == VB2_AGG_SUM_I64 || kKind == VB2_AGG_COUNT_MERGE) iv[r] = input_as_i64(u, jj[r]);
      if (row_valid) ok[r] = ok[r] && bit_at(row_valid, ii[r]);
      if (kKind != 0) {
        if (u.mask) ok[r] = ok[r] && bit_at(u.mask, ii[r]);
        if (u.nulls) ok[r] = ok[r] && bit_at(u.nulls, ii[r]);
        if (u.base_nulls) ok[r] = ok[r] && bit_at(u.base_nulls, jj[r]);
      }
      if (ok[r] && key[r] >= static_cast<uint64_t>(tab.capacity)) { bad = true; ok[r] = false; }
    }
#pragma unroll
    for (int r = 0; r < kU; ++r) {
      const int32_t g = ok[r] ? static_cast<int32_t>(key[r]) : -1;
#pragma unroll
      for (int k = 0; k < kTinyG; ++k) {
        if (g == k) {
          cnt[k] += 1;
          if (kKind == VB2_AGG_SUM_F64) fs[k] = __dadd_rn(fs[k], fv[r]);
          if (kKind == VB2_AGG_SUM_I64 || kKind == VB2_AGG_COUNT_MERGE) ovf |= add_overflow_i64(is[k], iv[r], &is[k]);
        }
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
    int64_t v = is[k];  // integer partials: detect overflow while combining
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const int64_t other = __shfl_xor_sync(0xffffffffu, v, o);
      ovf |= add_overflow_i64(v, other, &v);
    }
    const int64_t c = warp_sum(cnt[k]);
    if (lane == 0) { sf[warp][k] = f; si[warp][k] = v; sc[warp][k] = c; }
  }
  if (__any_sync(0xffffffffu, ovf) && lane == 0) atomicCAS(error_flag, 0, 1);
  if (__any_sync(0xffffffffu, bad) && lane == 0) atomicCAS(error_flag, 0, 100);
  __syncthreads();
  if (threadIdx.x < kTinyG && threadIdx.x < tab.capacity) {
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
      uint64_t* row = tab.rows + static_cast<int64_t>(k) * tab.row_words;
      if (kKind == 0) {
        atomicAdd(reinterpret_cast<unsigned long long*>(row), static_cast<unsigned long long>(c));
      } else {
        uint64_t* acc = row + u.acc_word;
        if (kKind == VB2_AGG_SUM_F64) atomicAdd(reinterpret_cast<double*>(acc), f);
        if (kKind == VB2_AGG_SUM_I64 || kKind == VB2_AGG_COUNT_MERGE) {
          const int64_t old = static_cast<int64_t>(atomicAdd(reinterpret_cast<unsigned long long*>(acc), static_cast<unsigned long long>(v)));
          int64_t r;
          o2 |= add_overflow_i64(old, v, &r);
        }
        if (kKind == VB2_AGG_COUNT) atomicAdd(reinterpret_cast<unsigned long long*>(acc), static_cast<unsigned long long>(c));
        if (u.nonnull_word >= 0 && kKind != VB2_AGG_COUNT) atomicAdd(reinterpret_cast<unsigned long long*>(row + u.nonnull_word), static_cast<unsigned long long>(c));
      }
    }
    if (o2) atomicCAS(error_flag, 0, 1);
  }
}

__global__ void table_init_kernel(uint64_t* __restrict__ rows, int64_t total_words, int32_t w, const __grid_constant__ vb2_group_row_init init) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total_words; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    rows[i] = init.words[i % w];
}

__device__ __forceinline__ bool row_occupied(const vb2_group_table& t, int64_t r) {
  const uint64_t w0 = t.rows[r * t.row_words];
  return t.hash_mode ? w0 != VB2_EMPTY_KEY : w0 != 0;
}
__device__ __forceinline__ uint64_t key_of_slot(const vb2_group_table& t, int64_t slot) {
  return t.hash_mode ? t.rows[slot * t.row_words] : static_cast<uint64_t>(slot);
}

__global__ void occupied_bits_kernel(const __grid_constant__ vb2_group_table t, uint32_t* __restrict__ bits) {
  const int64_t nwords = (t.capacity + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t s = (w << 5) + lane;
    const bool occ = s < t.capacity && row_occupied(t, s);
    const unsigned word = __ballot_sync(0xffffffffu, occ);
    if (lane == 0) bits[w] = word;
  }
}

// word <- value in every occupied row (a non-null counter that starts being tracked late).
__global__ void set_word_kernel(const __grid_constant__ vb2_group_table t, int32_t word, uint64_t value) {
  for (int64_t r = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; r < t.capacity; r += static_cast<int64_t>(gridDim.x) * blockDim.x)
    if (row_occupied(t, r)) t.rows[r * t.row_words + word] = value;
}

// Keys of occupied slots back to per-column values: id_k = (key / mult_k) % range_k.
__global__ void group_keys_kernel(const __grid_constant__ vb2_group_table t, const int32_t* __restrict__ slots, int64_t n,
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
      const uint64_t key = key_of_slot(t, slots[i]);
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

// Re-encodes the keys of occupied slots for a new layout (value ranges grew): decodes the
// per-column ids with the old (mult, range), re-bases them on the new mins and packs again.
struct RekeyArgs {
  int n;
  int64_t old_min[kMaxNormCols], new_min[kMaxNormCols];
  uint64_t old_mult[kMaxNormCols], old_range[kMaxNormCols], new_mult[kMaxNormCols];
  int old_null_reserved[kMaxNormCols];
};
__global__ void rekey_kernel(const __grid_constant__ vb2_group_table t, const int32_t* __restrict__ slots, int64_t n,
                             const __grid_constant__ RekeyArgs a, uint64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint64_t key = key_of_slot(t, slots[i]);
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

// Rehash / relayout: group i of the old table (slot slots[i]) moves to the row of new_keys[i] in
// the new table. Keys are distinct, so after the insert the row is owned by this thread.
__global__ void group_move_kernel(const __grid_constant__ vb2_group_table from, const int32_t* __restrict__ slots,
                                  const uint64_t* __restrict__ new_keys, int64_t n, const __grid_constant__ vb2_group_table to,
                                  int64_t* __restrict__ num_groups, int32_t* __restrict__ error_flag) {
  const TableView t = view_of(to);
  int64_t fresh = 0;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int64_t slot = find_or_insert(t, new_keys[i], fresh);
    if (slot < 0) { atomicCAS(error_flag, 0, 100); continue; }
    const uint64_t* src = from.rows + static_cast<int64_t>(slots[i]) * from.row_words;
    uint64_t* dst = to.rows + slot * to.row_words;
    if (!to.hash_mode) dst[0] = 1;
    for (int w = 1; w < to.row_words; ++w) dst[w] = src[w];
  }
  fresh = warp_sum(fresh);
  if ((threadIdx.x & 31) == 0 && fresh && num_groups) atomicAdd(reinterpret_cast<unsigned long long*>(num_groups), static_cast<unsigned long long>(fresh));
}

__global__ void group_gather_kernel(const __grid_constant__ vb2_group_table t, const int32_t* __restrict__ slots, int64_t n, int32_t word,
                                    uint64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = t.rows[static_cast<int64_t>(slots[i]) * t.row_words + word];
}
__global__ void group_valid_kernel(const __grid_constant__ vb2_group_table t, const int32_t* __restrict__ slots, int64_t n, int32_t word,
                                   uint32_t* __restrict__ out) {
  const int64_t nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t k = (w << 5) + lane;
    const bool b = k < n && static_cast<int64_t>(t.rows[static_cast<int64_t>(slots[k]) * t.row_words + word]) > 0;
    const unsigned bits = __ballot_sync(0xffffffffu, b);
    if (lane == 0) out[w] = bits;
  }
}
__global__ void group_avg_kernel(const __grid_constant__ vb2_group_table t, const int32_t* __restrict__ slots, int64_t n, int32_t sum_word,
                                 int32_t count_word, double* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const uint64_t* row = t.rows + static_cast<int64_t>(slots[i]) * t.row_words;
    const int64_t c = static_cast<int64_t>(row[count_word]);
    // functions/lib/aggregates/AverageAggregateBase.h:86-107
    out[i] = c > 0 ? __ddiv_rn(__longlong_as_double(static_cast<long long>(row[sum_word])), static_cast<double>(c)) : 0.0;
  }
}

// Every output column of an aggregation in one launch. Small tables are compacted by the kernel
// itself (single block: ballot + running prefix over the slots in ascending order), so the host
// learns the row count from the same copy that brings the result over.
struct ExtractArgs {
  vb2_extract_col c[VB2_EXTRACT_MAX_COLS];
  int n;
};
__global__ void __launch_bounds__(1024)
group_extract_kernel(const __grid_constant__ vb2_group_table t, const int32_t* __restrict__ slots, int64_t n, int32_t* __restrict__ scratch,
                     const __grid_constant__ ExtractArgs a, int64_t* __restrict__ header, const int32_t* __restrict__ error_flag) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (slots == nullptr) {
    __shared__ int warp_tot[32];
    int64_t base = 0;
    const int nwarps = blockDim.x >> 5;
    for (int64_t r0 = 0; r0 < t.capacity; r0 += blockDim.x) {
      const int64_t r = r0 + threadIdx.x;
      const bool occ = r < t.capacity && row_occupied(t, r);
      const unsigned m = __ballot_sync(0xffffffffu, occ);
      if (lane == 0) warp_tot[warp] = __popc(m);
      __syncthreads();
      int before = 0, total = 0;
      for (int w = 0; w < nwarps; ++w) {
        const int c = warp_tot[w];
        if (w < warp) before += c;
        total += c;
      }
      if (occ) scratch[base + before + __popc(m & ((1u << lane) - 1u))] = static_cast<int32_t>(r);
      base += total;
      __syncthreads();
    }
    n = base;
    slots = scratch;
    if (threadIdx.x == 0 && header) header[0] = n;
    __syncthreads();
  } else if (blockIdx.x == 0 && threadIdx.x == 0 && header) {
    header[0] = n;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0 && header) header[1] = error_flag ? *error_flag : 0;
  const int64_t nwords = (n + 31) >> 5;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps_grid = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps_grid) {
    const int64_t i = (w << 5) + lane;
    const bool live = i < n;
    const int64_t slot = live ? slots[i] : 0;
    const uint64_t* row = t.rows + slot * t.row_words;
    const uint64_t key = t.hash_mode ? row[0] : static_cast<uint64_t>(slot);
    const bool last_even_word = w == nwords - 1 && (w & 1) == 0;  // the upper half of the last 64-bit word stays clear
    for (int k = 0; k < a.n; ++k) {
      const vb2_extract_col& c = a.c[k];
      bool valid = live;
      if (c.kind == VB2_EXTRACT_KEY) {
        const uint64_t id = (key / c.mult) % c.range;
        valid = live && !(c.null_reserved && id == 0);
        const int64_t v = valid ? static_cast<int64_t>(id) - 1 + c.min : 0;
        if (c.type == VB2_BOOLEAN) {
          const unsigned bitsv = __ballot_sync(0xffffffffu, live && v != 0);
          if (lane == 0) {
            reinterpret_cast<uint32_t*>(c.values)[w] = bitsv;
            if (last_even_word) reinterpret_cast<uint32_t*>(c.values)[w + 1] = 0;
          }
        } else if (live) {
          if (c.type == VB2_INTEGER) reinterpret_cast<int32_t*>(c.values)[i] = static_cast<int32_t>(v);
          else reinterpret_cast<int64_t*>(c.values)[i] = v;
        }
      } else if (c.kind == VB2_EXTRACT_KEYWORD) {
        const uint64_t word = row[c.word];
        valid = live && !((row[c.count_word] >> c.null_reserved) & 1u);
        if (c.type == VB2_BOOLEAN) {
          const unsigned bitsv = __ballot_sync(0xffffffffu, valid && word != 0);
          if (lane == 0) {
            reinterpret_cast<uint32_t*>(c.values)[w] = bitsv;
            if (last_even_word) reinterpret_cast<uint32_t*>(c.values)[w + 1] = 0;
          }
        } else if (live) {
          // c.min: offset for id-coded keys (VARCHAR dictionary index = global id - 1); 0 otherwise
          if (c.type == VB2_INTEGER) reinterpret_cast<int32_t*>(c.values)[i] = valid ? static_cast<int32_t>(static_cast<int64_t>(word) + c.min) : 0;
          else if (c.type == VB2_DOUBLE) reinterpret_cast<uint64_t*>(c.values)[i] = valid ? word : 0;  // the (canonical) bits
          else reinterpret_cast<int64_t*>(c.values)[i] = valid ? static_cast<int64_t>(word) + c.min : 0;
        }
      } else {
        const int64_t cnt = c.count_word >= 0 ? static_cast<int64_t>(row[c.count_word]) : 1;
        valid = live && cnt > 0;
        if (live) {
          const uint64_t word = row[c.word];
          if (c.kind == VB2_EXTRACT_WORD) reinterpret_cast<uint64_t*>(c.values)[i] = word;
          else if (c.kind == VB2_EXTRACT_WORD_I32) reinterpret_cast<int32_t*>(c.values)[i] = static_cast<int32_t>(static_cast<int64_t>(word));
          else reinterpret_cast<double*>(c.values)[i] = cnt > 0 ? __ddiv_rn(__longlong_as_double(static_cast<long long>(word)), static_cast<double>(cnt)) : 0.0;
        }
      }
      if (c.valid) {
        const unsigned bitsv = __ballot_sync(0xffffffffu, valid);
        if (lane == 0) {
          reinterpret_cast<uint32_t*>(c.valid)[w] = bitsv;
          if (last_even_word) reinterpret_cast<uint32_t*>(c.valid)[w + 1] = 0;
        }
      }
    }
  }
}

// Partial results of a fused scan (sums[g * nproj + p], counts[g]) added into the table rows of
// an array-mode table (group g = row g): one thread per (group, target word).
struct MergeArgs {
  int32_t word[2 * kMaxAggs + 1];
  int32_t proj[2 * kMaxAggs + 1];  // >= 0: += sums[g * nproj + proj] (double); -1: += counts[g] (int64)
  int n;
};
__global__ void merge_partials_kernel(const __grid_constant__ vb2_group_table t, const double* __restrict__ sums, const int64_t* __restrict__ counts,
                                      int32_t ngroups, int32_t nproj, const __grid_constant__ MergeArgs m) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= ngroups * m.n) return;
  const int g = i / m.n, k = i % m.n;
  const int64_t c = counts[g];
  if (c == 0) return;  // no row of the batch reached this group
  uint64_t* p = t.rows + static_cast<int64_t>(g) * t.row_words + m.word[k];
  if (m.proj[k] >= 0) {
    const double d = __dadd_rn(__longlong_as_double(static_cast<long long>(*p)), sums[static_cast<int64_t>(g) * nproj + m.proj[k]]);
    *p = static_cast<uint64_t>(__double_as_longlong(d));
  } else {
    *p = static_cast<uint64_t>(static_cast<int64_t>(*p) + c);
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
  normalize_keys_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(a, sel, n, keys_out, reinterpret_cast<uint32_t*>(valid_out));
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_dictionary_codes(const vb2_column* col, int64_t n, int32_t* codes, uint8_t* valid, void* stream) {
  if (!col || col->encoding == VB2_FLAT) return fail_msg(VB2_ERR_INVALID, "dictionary_codes: dictionary or constant column expected");
  if (n <= 0) return VB2_OK;
  dictionary_codes_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*col, n, codes, valid);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_join_build_array_direct(int32_t* head, int32_t* next, int64_t capacity, const vb2_column* key, int64_t lo, int64_t n,
                                 int32_t* flags, void* stream) {
  if (!head || !next || !key || capacity <= 0) return fail_msg(VB2_ERR_INVALID, "join_build_array_direct: bad arguments");
  if (key->type != VB2_BIGINT && key->type != VB2_INTEGER && key->type != VB2_BOOLEAN) return fail_msg(VB2_ERR_INVALID, "join_build_array_direct: integer-typed key expected");
  if (n <= 0) return VB2_OK;
  join_build_array_direct_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(head, next, capacity, *key, lo, n, flags);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_column_minmax(const vb2_column* col, int64_t rows, int64_t* out3, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  minmax_init_kernel<<<vb2::counted(1), 1, 0, st>>>(out3);
  if (rows > 0 && col->encoding == VB2_FLAT && !col->nulls && col->type == VB2_BIGINT && (reinterpret_cast<uintptr_t>(col->values) & 15) == 0)
    minmax_flat_i64_kernel<<<vb2::counted(grid_for(rows / 2 + 1, 256)), 256, 0, st>>>(static_cast<const int64_t*>(col->values), rows, out3);
  else if (rows > 0) minmax_kernel<<<vb2::counted(grid_for(rows, 256)), 256, 0, st>>>(*col, rows, out3);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

static int check_table(const vb2_group_table* t, const char* who) {
  if (!t || !t->rows || t->capacity <= 0 || t->row_words < 1 || t->row_words > VB2_MAX_ROW_WORDS) return fail_msg(VB2_ERR_INVALID, who);
  if (t->hash_mode && (t->capacity & (t->capacity - 1))) return fail_msg(VB2_ERR_INVALID, "group table: hash-mode capacity must be a power of two");
  if (t->capacity > (1ll << 31)) return fail_msg(VB2_ERR_UNSUPPORTED, "group table: capacity above 2^31 rows");
  return VB2_OK;
}

int vb2k_group_table_init(const vb2_group_table* t, const uint64_t* row_init, void* stream) {
  if (int rc = check_table(t, "group_table_init: bad table")) return rc;
  vb2_group_row_init init;
  for (int i = 0; i < t->row_words; ++i) init.words[i] = row_init[i];
  const int64_t total = t->capacity * t->row_words;
  table_init_kernel<<<vb2::counted(grid_for(total, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(t->rows, total, t->row_words, init);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_update(const vb2_group_table* t, const uint64_t* row_keys, const uint64_t* row_valid, int64_t n,
                      const vb2_agg_update* aggs, int32_t naggs, int64_t* num_groups, int32_t* error_flag, void* stream) {
  if (int rc = check_table(t, "group_update: bad table")) return rc;
  if (naggs < 0 || naggs > kMaxAggs) return fail_msg(VB2_ERR_UNSUPPORTED, "group_update: at most 16 aggregates per call");
  if (n <= 0) return VB2_OK;
  for (int i = 0; i < naggs; ++i)
    if (aggs[i].acc_word < 1 || aggs[i].acc_word >= t->row_words || aggs[i].nonnull_word >= t->row_words)
      return fail_msg(VB2_ERR_INVALID, "group_update: accumulator word outside the row");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  AggArgs rest;
  rest.n = 0;
  // aggregates of one call must not share an accumulator word for the serial kernel's lane ownership
  bool distinct_words = true;
  for (int i = 0; i < naggs; ++i)
    for (int j = 0; j < i; ++j)
      if (aggs[i].acc_word == aggs[j].acc_word || (aggs[i].nonnull_word >= 0 && (aggs[i].nonnull_word == aggs[j].nonnull_word || aggs[i].nonnull_word == aggs[j].acc_word)) ||
          (aggs[j].nonnull_word >= 0 && aggs[j].nonnull_word == aggs[i].acc_word))
        distinct_words = false;
  const bool tiny_table = !t->hash_mode && t->capacity <= kTinyG;
  if (n <= kSerialRows && distinct_words) {
    for (int i = 0; i < naggs; ++i) rest.a[rest.n++] = aggs[i];
    group_update_serial_kernel<<<vb2::counted(1), 32, 0, st>>>(*t, row_keys, row_valid, n, rest, num_groups, error_flag);
  } else if (tiny_table && n > kAtomicRows) {
    const unsigned grid = grid_for(n, 256, 4);
    vb2_agg_update none{};
    group_update_tiny_kernel<0><<<vb2::counted(grid), 256, 0, st>>>(*t, row_keys, row_valid, n, none, error_flag);
    for (int i = 0; i < naggs; ++i) {
      const vb2_agg_update& u = aggs[i];
      switch (u.kind) {
        case VB2_AGG_SUM_F64: group_update_tiny_kernel<VB2_AGG_SUM_F64><<<vb2::counted(grid), 256, 0, st>>>(*t, row_keys, row_valid, n, u, error_flag); break;
        case VB2_AGG_SUM_I64: group_update_tiny_kernel<VB2_AGG_SUM_I64><<<vb2::counted(grid), 256, 0, st>>>(*t, row_keys, row_valid, n, u, error_flag); break;
        case VB2_AGG_COUNT_MERGE: group_update_tiny_kernel<VB2_AGG_COUNT_MERGE><<<vb2::counted(grid), 256, 0, st>>>(*t, row_keys, row_valid, n, u, error_flag); break;
        case VB2_AGG_COUNT: group_update_tiny_kernel<VB2_AGG_COUNT><<<vb2::counted(grid), 256, 0, st>>>(*t, row_keys, row_valid, n, u, error_flag); break;
        default: rest.a[rest.n++] = u; break;  // min / max: idempotent atomics, no same-address accumulation chain
      }
    }
    if (rest.n) group_update_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(*t, row_keys, row_valid, n, rest, nullptr, error_flag);
  } else if (!t->hash_mode && row_keys && distinct_words && n > kAtomicRows && t->capacity > kTinyG &&
             t->capacity * t->row_words * 8 <= kSmemTableBytes) {
    // low-cardinality array table: block-private shared-memory copies, merged once per block
    for (int i = 0; i < naggs; ++i) rest.a[rest.n++] = aggs[i];
    const size_t smem = static_cast<size_t>(t->capacity) * t->row_words * 8;
    group_update_smem_kernel<<<vb2::counted(grid_for(n, 256, 8)), 256, smem, st>>>(*t, row_keys, row_valid, n, rest, error_flag);
  } else {
    for (int i = 0; i < naggs; ++i) rest.a[rest.n++] = aggs[i];
    group_update_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(*t, row_keys, row_valid, n, rest, num_groups, error_flag);
  }
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_update_partitioned(const vb2_group_table* t, const uint64_t* row_keys, const int64_t* part_start, int32_t nparts, int64_t n,
                                  const vb2_agg_update* aggs, int32_t naggs, int64_t* num_groups, int32_t* error_flag, uint32_t* barrier_word,
                                  void* stream) {
  if (int rc = check_table(t, "group_update_partitioned: bad table")) return rc;
  if (t->hash_mode != 1 || nparts < 1 || nparts > 256 || t->capacity < 65536 || (t->capacity % nparts) != 0)
    return fail_msg(VB2_ERR_UNSUPPORTED, "group_update_partitioned: hash-mode table of at least 65536 rows expected");
  if (naggs < 0 || naggs > kMaxAggs) return fail_msg(VB2_ERR_UNSUPPORTED, "group_update_partitioned: at most 16 aggregates per call");
  if (n <= 0) return VB2_OK;
  AggArgs a;
  a.n = naggs;
  for (int i = 0; i < naggs; ++i) {
    if (aggs[i].acc_word < 1 || aggs[i].acc_word >= t->row_words || aggs[i].nonnull_word >= t->row_words)
      return fail_msg(VB2_ERR_INVALID, "group_update_partitioned: accumulator word outside the row");
    a.a[i] = aggs[i];
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static int blocks_per_sm = 0;
  if (blocks_per_sm == 0) {
    int b = 0;
    VB2_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, group_update_partitioned_kernel, 256, 0));
    blocks_per_sm = b < 1 ? 1 : (b > 4 ? 4 : b);
  }
  const unsigned grid = static_cast<unsigned>(device_sm_count() * blocks_per_sm);
  VB2_CUDA_OK(cudaMemsetAsync(barrier_word, 0, 4, st));
  vb2_group_table tab = *t;
  void* params[] = {&tab, &row_keys, &part_start, &nparts, &a, &num_groups, &error_flag, &barrier_word};
  note_launch();
  // cooperative launch: every block is resident, which the grid barrier relies on
  VB2_CUDA_OK(cudaLaunchCooperativeKernel(reinterpret_cast<const void*>(group_update_partitioned_kernel), dim3(grid), dim3(256), params, 0, st));
  return VB2_OK;
}

size_t vb2k_group_occupied_workspace(int64_t capacity) {
  return static_cast<size_t>((capacity + 63) >> 6) * 8 + vb2k_bits_to_indices_workspace(capacity);
}

int vb2k_group_occupied(const vb2_group_table* t, int32_t* slot_list, int64_t* count, void* workspace, size_t workspace_bytes, void* stream) {
  if (int rc = check_table(t, "group_occupied: bad table")) return rc;
  // workspace = occupancy bitmap (capacity bits, 8-byte aligned) followed by the compaction scratch
  const size_t bitmap_bytes = static_cast<size_t>((t->capacity + 63) >> 6) * 8;
  if (workspace_bytes < vb2k_group_occupied_workspace(t->capacity)) return fail_msg(VB2_ERR_INVALID, "group_occupied: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  uint64_t* bits = reinterpret_cast<uint64_t*>(workspace);
  VB2_CUDA_OK(cudaMemsetAsync(bits, 0, bitmap_bytes, st));
  occupied_bits_kernel<<<vb2::counted(grid_for(t->capacity, 256)), 256, 0, st>>>(*t, reinterpret_cast<uint32_t*>(bits));
  VB2_CUDA_OK(cudaGetLastError());
  return vb2k_bits_to_indices(bits, t->capacity, slot_list, count, reinterpret_cast<char*>(workspace) + bitmap_bytes,
                              workspace_bytes - bitmap_bytes, stream);
}

int vb2k_group_set_word(const vb2_group_table* t, int32_t word, uint64_t value, void* stream) {
  if (int rc = check_table(t, "group_set_word: bad table")) return rc;
  if (word < 1 || word >= t->row_words) return fail_msg(VB2_ERR_INVALID, "group_set_word: word outside the row");
  set_word_kernel<<<vb2::counted(grid_for(t->capacity, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*t, word, value);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_keys(const vb2_group_table* t, const int32_t* slots, int64_t n, int64_t min, uint64_t mult, uint64_t range,
                    int32_t null_reserved, int32_t type, void* values, uint64_t* valid, void* stream) {
  if (int rc = check_table(t, "group_keys: bad table")) return rc;
  if (n <= 0) return VB2_OK;
  group_keys_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*t, slots, n, min, mult, range, null_reserved, type, values,
                                                                                      reinterpret_cast<uint32_t*>(valid));
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_rekey(const vb2_group_table* t, const int32_t* slots, int64_t n, int32_t ncols, const int64_t* old_mins,
                     const uint64_t* old_mults, const uint64_t* old_ranges, const int32_t* old_null_reserved, const int64_t* new_mins,
                     const uint64_t* new_mults, uint64_t* keys_out, void* stream) {
  if (int rc = check_table(t, "group_rekey: bad table")) return rc;
  if (ncols < 1 || ncols > kMaxNormCols) return fail_msg(VB2_ERR_UNSUPPORTED, "group_rekey: 1..4 key columns");
  if (n <= 0) return VB2_OK;
  RekeyArgs a;
  a.n = ncols;
  for (int i = 0; i < ncols; ++i) {
    a.old_min[i] = old_mins[i]; a.new_min[i] = new_mins[i];
    a.old_mult[i] = old_mults[i]; a.old_range[i] = old_ranges[i]; a.new_mult[i] = new_mults[i];
    a.old_null_reserved[i] = old_null_reserved ? old_null_reserved[i] : 1;
  }
  rekey_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*t, slots, n, a, keys_out);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_move(const vb2_group_table* from, const int32_t* slots, const uint64_t* new_keys, int64_t n, const vb2_group_table* to,
                    int64_t* num_groups, int32_t* error_flag, void* stream) {
  if (int rc = check_table(from, "group_move: bad source table")) return rc;
  if (int rc = check_table(to, "group_move: bad target table")) return rc;
  if (from->row_words != to->row_words) return fail_msg(VB2_ERR_INVALID, "group_move: row layouts differ");
  if (n <= 0) return VB2_OK;
  group_move_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*from, slots, new_keys, n, *to, num_groups, error_flag);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_gather(const vb2_group_table* t, const int32_t* slots, int64_t n, int32_t word, void* out, void* stream) {
  if (int rc = check_table(t, "group_gather: bad table")) return rc;
  if (word < 0 || word >= t->row_words) return fail_msg(VB2_ERR_INVALID, "group_gather: word outside the row");
  if (n <= 0) return VB2_OK;
  group_gather_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*t, slots, n, word, reinterpret_cast<uint64_t*>(out));
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_valid(const vb2_group_table* t, const int32_t* slots, int64_t n, int32_t count_word, uint64_t* valid, void* stream) {
  if (int rc = check_table(t, "group_valid: bad table")) return rc;
  if (count_word < 0 || count_word >= t->row_words) return fail_msg(VB2_ERR_INVALID, "group_valid: word outside the row");
  if (n <= 0) return VB2_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VB2_CUDA_OK(cudaMemsetAsync(valid + ((n + 63) >> 6) - 1, 0, 8, st));
  group_valid_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(*t, slots, n, count_word, reinterpret_cast<uint32_t*>(valid));
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_avg(const vb2_group_table* t, const int32_t* slots, int64_t n, int32_t sum_word, int32_t count_word, double* out, void* stream) {
  if (int rc = check_table(t, "group_avg: bad table")) return rc;
  if (n <= 0) return VB2_OK;
  group_avg_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*t, slots, n, sum_word, count_word, out);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

static int check_keyed(const vb2_group_table* t, int32_t nkeys, const char* who) {
  if (int rc = check_table(t, who)) return rc;
  if (t->hash_mode != VB2_GROUP_KEYED || nkeys < 1 || nkeys > VB2_KEYED_MAX_KEYS || t->row_words < nkeys + 2) return fail_msg(VB2_ERR_INVALID, who);
  return VB2_OK;
}

int vb2k_group_update_keyed(const vb2_group_table* t, const vb2_column* keys, int32_t nkeys, int64_t n, const vb2_agg_update* aggs, int32_t naggs,
                            int64_t* num_groups, int32_t* error_flag, void* stream) {
  if (int rc = check_keyed(t, nkeys, "group_update_keyed: bad table")) return rc;
  if (naggs < 0 || naggs > kMaxAggs) return fail_msg(VB2_ERR_UNSUPPORTED, "group_update_keyed: at most 16 aggregates per call");
  if (n <= 0) return VB2_OK;
  KeyedCols kc;
  kc.n = nkeys;
  for (int k = 0; k < nkeys; ++k) {
    if (keys[k].type == VB2_VARCHAR) return fail_msg(VB2_ERR_INVALID, "group_update_keyed: VARCHAR keys arrive as id columns");
    kc.c[k] = keys[k];
  }
  AggArgs a;
  a.n = naggs;
  for (int i = 0; i < naggs; ++i) {
    if (aggs[i].acc_word < nkeys + 2 || aggs[i].acc_word >= t->row_words || aggs[i].nonnull_word >= t->row_words)
      return fail_msg(VB2_ERR_INVALID, "group_update_keyed: accumulator word outside the row");
    a.a[i] = aggs[i];
  }
  group_update_keyed_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*t, kc, n, a, num_groups, error_flag);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_keyed_key_ids(const vb2_group_table* t, const vb2_column* keys, int32_t nkeys, int64_t n, int32_t insert, uint64_t* ids, uint64_t* valid,
                       int64_t* num_groups, int32_t* error_flag, void* stream) {
  if (int rc = check_keyed(t, nkeys, "keyed_key_ids: bad table")) return rc;
  if (n <= 0) return VB2_OK;
  KeyedCols kc{};
  kc.n = nkeys;
  for (int k = 0; k < nkeys; ++k) {
    if (keys[k].type == VB2_VARCHAR) return fail_msg(VB2_ERR_INVALID, "keyed_key_ids: VARCHAR keys arrive as id columns");
    kc.c[k] = keys[k];
  }
  keyed_key_ids_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*t, kc, n, insert, ids, reinterpret_cast<uint32_t*>(valid),
                                                                                                    num_groups, error_flag);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_move_keyed(const vb2_group_table* from, const int32_t* slots, int64_t n, int32_t nkeys, const vb2_group_table* to, int64_t* num_groups,
                          int32_t* error_flag, void* stream) {
  if (int rc = check_keyed(from, nkeys, "group_move_keyed: bad source table")) return rc;
  if (int rc = check_keyed(to, nkeys, "group_move_keyed: bad target table")) return rc;
  if (from->row_words != to->row_words) return fail_msg(VB2_ERR_INVALID, "group_move_keyed: row layouts differ");
  if (n <= 0) return VB2_OK;
  DecodeArgs d{};
  d.n = nkeys;
  d.from_keyed = 1;
  group_move_keyed_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*from, slots, n, d, *to, num_groups, error_flag);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_move_to_keyed(const vb2_group_table* from, const int32_t* slots, int64_t n, int32_t nkeys, const int64_t* mins, const uint64_t* mults,
                             const uint64_t* ranges, const int32_t* null_reserved, int32_t word_shift, const vb2_group_table* to,
                             int64_t* num_groups, int32_t* error_flag, void* stream) {
  if (int rc = check_table(from, "group_move_to_keyed: bad source table")) return rc;
  if (int rc = check_keyed(to, nkeys, "group_move_to_keyed: bad target table")) return rc;
  if (from->hash_mode == VB2_GROUP_KEYED || word_shift != nkeys + 1) return fail_msg(VB2_ERR_INVALID, "group_move_to_keyed: bad layouts");
  if (n <= 0) return VB2_OK;
  DecodeArgs d{};
  d.n = nkeys;
  d.word_shift = word_shift;
  for (int k = 0; k < nkeys; ++k) {
    d.mins[k] = mins[k];
    d.mults[k] = mults[k] ? mults[k] : 1;
    d.ranges[k] = ranges[k] ? ranges[k] : 1;
    d.null_reserved[k] = null_reserved ? null_reserved[k] : 0;
  }
  group_move_keyed_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(*from, slots, n, d, *to, num_groups, error_flag);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_extract(const vb2_group_table* t, const int32_t* slots, int64_t n, int32_t* scratch_slots, const vb2_extract_col* cols,
                       int32_t ncols, int64_t* header, const int32_t* error_flag, void* stream) {
  if (int rc = check_table(t, "group_extract: bad table")) return rc;
  if (ncols < 0 || ncols > VB2_EXTRACT_MAX_COLS) return fail_msg(VB2_ERR_UNSUPPORTED, "group_extract: too many output columns");
  if (!slots && (t->capacity > VB2_EXTRACT_SMALL_CAPACITY || !scratch_slots)) return fail_msg(VB2_ERR_INVALID, "group_extract: slot list required for large tables");
  ExtractArgs a;
  a.n = ncols;
  for (int i = 0; i < ncols; ++i) {
    const vb2_extract_col& c = cols[i];
    if (c.kind < VB2_EXTRACT_KEY || c.kind > VB2_EXTRACT_KEYWORD || !c.values) return fail_msg(VB2_ERR_INVALID, "group_extract: bad column");
    if (c.kind != VB2_EXTRACT_KEY && (c.word < 0 || c.word >= t->row_words || c.count_word >= t->row_words)) return fail_msg(VB2_ERR_INVALID, "group_extract: word outside the row");
    if (c.kind == VB2_EXTRACT_KEY && (c.mult == 0 || c.range == 0)) return fail_msg(VB2_ERR_INVALID, "group_extract: bad key layout");
    a.c[i] = c;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!slots) {
    group_extract_kernel<<<vb2::counted(1), 1024, 0, st>>>(*t, nullptr, 0, scratch_slots, a, header, error_flag);
  } else {
    if (n <= 0) return VB2_OK;
    group_extract_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(*t, slots, n, nullptr, a, header, error_flag);
  }
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_group_merge_partials(const vb2_group_table* t, const double* sums, const int64_t* counts, int32_t ngroups, int32_t nproj,
                              const int32_t* target_words, const int32_t* target_projs, int32_t ntargets, void* stream) {
  if (int rc = check_table(t, "group_merge_partials: bad table")) return rc;
  if (t->hash_mode || ngroups > t->capacity) return fail_msg(VB2_ERR_INVALID, "group_merge_partials: array-mode table of >= ngroups rows expected");
  if (ntargets < 0 || ntargets > 2 * kMaxAggs + 1) return fail_msg(VB2_ERR_UNSUPPORTED, "group_merge_partials: too many target words");
  if (ngroups <= 0 || ntargets == 0) return VB2_OK;
  MergeArgs m;
  m.n = ntargets;
  for (int i = 0; i < ntargets; ++i) {
    if (target_words[i] < 0 || target_words[i] >= t->row_words || target_projs[i] >= nproj) return fail_msg(VB2_ERR_INVALID, "group_merge_partials: bad target");
    m.word[i] = target_words[i];
    m.proj[i] = target_projs[i];
  }
  const int total = ngroups * ntargets;
  merge_partials_kernel<<<vb2::counted((total + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(*t, sums, counts, ngroups, nproj, m);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

}  // extern "C"
