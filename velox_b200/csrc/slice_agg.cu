// High-cardinality GROUP BY entirely on chip: two radix-partition passes cut the buffered input into
// slices whose distinct keys fit one CTA's shared-memory hash table, every slice is then aggregated in
// shared memory and its groups are appended to a compact array of group rows.
//
// Replaces, for inputs with millions of groups, the find-or-insert of HashTable::groupProbe /
// ProbeState::fullProbe + the accumulator scatter of SimpleNumericAggregate (velox/exec/HashTable.cpp:
// 470-519,138; functions/lib/aggregates/SimpleNumericAggregate.h:94-150): on the CPU that loop is
// cache / TLB-miss bound and the reference hides it with a 4-way interleave and prefetches; on B200 a
// global table of 100 M x 32 B rows costs one random DRAM sector read-modify-write plus a page walk
// per input row (198 B of DRAM reads per row measured, profiles/r01_ncu_config5_group_update.json).
// Here DRAM only sees streaming passes:
//
//   level 1   rows -> 256 partitions by bits 63..56 of a Fibonacci hash   (histogram, scan, staged scatter)
//   level 2   every partition -> P2 <= 256 slices by the next bits          (same kernels, one launch each)
//   aggregate one CTA per slice: open-addressing table in shared memory (keys + accumulator words),
//             smem atomics, then the occupied slots become group rows [key | accumulators] in the
//             vb2_group_table row layout, appended at an atomically reserved offset
//
// The scatter stages each tile in shared memory in partition order, so global stores are runs of
// consecutive rows (full 32-byte sectors) — a direct 8-byte scatter makes DRAM read and write every
// sector twice. All loads of the input are coalesced and streaming.
// A HyperLogLog sketch filled by the level-1 histogram sizes level 2 and the output.
#include <cstdlib>

#include "common.cuh"

namespace vb2 {

namespace {

constexpr int kPT = 512;          // threads of the partition kernels
constexpr int kTile = 3584;       // rows per tile = 7 per thread (held in registers, then staged in shared memory in partition order)
constexpr int kP1 = 256;
constexpr int kMaxP = 256;
constexpr int kCols = VB2_SLICE_MAX_COLS;
constexpr int kHllBits = 12;
constexpr int kAggThreads = 512;
constexpr int kMaxOps = VB2_SLICE_MAX_OPS;

struct SliceKey {
  const uint64_t* norm;  // normalized keys, or NULL: raw BIGINT keys below
  const int64_t* raw;
  int64_t min;           // normalized key = raw - min + 1 (vb2k_normalize_keys with one column)
};
// Per tile: a base pointer and one subtrahend (normalized key = word - sub), so that the row loops index
// with 32-bit in-tile offsets and carry no per-row branch on the key form.
struct TileKeys {
  const uint64_t* p;
  uint64_t sub;
};
__device__ __forceinline__ TileKeys tile_keys(const SliceKey& k, int64_t begin) {
  return k.norm ? TileKeys{k.norm + begin, 0} : TileKeys{reinterpret_cast<const uint64_t*>(k.raw) + begin, static_cast<uint64_t>(k.min) - 1};
}
// Every row is hashed four times on its way (two histograms, two scatters): twang_mix64 costs ~40 32-bit
// instructions and made those passes issue-bound, so the slice path places rows with a Fibonacci hash —
// one 64-bit multiply. Bits 63..56 pick the level-1 partition, the next log2(P2) bits the slice, bits
// 46..34 the slot inside the slice's table. (Placement only: nothing of this hash leaves the kernels.)
__device__ __forceinline__ uint64_t slice_mix(uint64_t key) { return key * 0x9E3779B97F4A7C15ull; }

struct PartIO {
  SliceKey key;
  const uint64_t* cols_in[kCols];
  uint64_t* keys_out;
  uint64_t* cols_out[kCols];
  int ncols;
};
struct PartGeom {
  int64_t n;                  // rows of the input array
  const int64_t* seg_start;   // device [nseg + 1]; NULL: one segment [0, n)
  const int32_t* tile_start;  // device [nseg + 1]: first tile of every segment (NULL with seg_start)
  int nseg;
  int P, shift;               // digit = (slice_mix(key) >> shift) & (P - 1)
};

// Segment table of a level-2 pass in shared memory: every tile looks its segment up (a binary search of eight
// dependent loads), which from global memory cost about as much as streaming the tile itself.
struct SegTable {
  int64_t seg_start[kP1 + 1];
  int32_t tile_start[kP1 + 1];
};
__device__ __forceinline__ void load_seg_table(const PartGeom& g, SegTable& t) {
  if (!g.seg_start) return;
  for (int i = threadIdx.x; i <= g.nseg; i += blockDim.x) {
    t.seg_start[i] = g.seg_start[i];
    t.tile_start[i] = g.tile_start[i];
  }
  __syncthreads();
}
// rows [begin, end) and segment of a tile; false past the last tile (tiles are ordered: every later tile is past it too)
__device__ __forceinline__ bool tile_range(const PartGeom& g, const SegTable& t, int64_t tile, int& seg, int64_t& begin, int64_t& end) {
  if (!g.seg_start) {
    seg = 0;
    begin = tile * kTile;
    if (begin >= g.n) return false;
    end = begin + kTile < g.n ? begin + kTile : g.n;
    return true;
  }
  if (tile >= t.tile_start[g.nseg]) return false;
  int lo = 0, hi = g.nseg;  // tile_start[lo] <= tile < tile_start[hi]
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if (t.tile_start[mid] <= tile) lo = mid;
    else hi = mid;
  }
  seg = lo;
  begin = t.seg_start[seg] + (tile - t.tile_start[seg]) * kTile;
  const int64_t segEnd = t.seg_start[seg + 1];
  end = begin + kTile < segEnd ? begin + kTile : segEnd;
  return true;
}

// hist[seg * P + d] += rows of segment seg with digit d = (slice_mix(key) >> shift) & (P - 1). A block walks a contiguous range of tiles and
// flushes its shared histogram when the segment changes. hll (optional): HyperLogLog registers over
// the keys whose hash ends in 000 (the convention of radix_partition.cu: the host multiplies by 8).
__global__ void __launch_bounds__(kPT) part_hist_kernel(const __grid_constant__ SliceKey key, const __grid_constant__ PartGeom g, int64_t ntiles,
                                                        uint32_t* __restrict__ hist, int32_t* __restrict__ hll) {
  __shared__ uint32_t h[kMaxP];
  __shared__ int32_t regs[1 << kHllBits];
  // Sampled keys wait in a per-warp queue until 32 of them are there: the strong hash of the sketch then runs
  // with every lane busy (one row in eight is sampled: hashing in place would run its ~40 instructions for
  // nearly every warp iteration with four lanes active).
  __shared__ uint64_t hq[kPT / 32][64];
  __shared__ SegTable segs;
  load_seg_table(g, segs);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  int qn = 0;  // warp-uniform
  auto sketch = [&](uint64_t k) {
    const uint64_t hash = twang_mix64(k);
    const uint32_t idx = static_cast<uint32_t>(hash >> 3) & ((1u << kHllBits) - 1u);
    const uint64_t rest = (hash >> (3 + kHllBits)) | (1ull << (56 - 3 - kHllBits));
    atomicMax(&regs[idx], __ffsll(static_cast<long long>(rest)));
  };
  if (hll)
    for (int i = threadIdx.x; i < (1 << kHllBits); i += kPT) regs[i] = 0;
  if (threadIdx.x < kMaxP) h[threadIdx.x] = 0;
  __syncthreads();
  const int64_t per = (ntiles + gridDim.x - 1) / gridDim.x;
  const int64_t t0 = blockIdx.x * per, t1 = t0 + per < ntiles ? t0 + per : ntiles;
  int segNow = -1;
  for (int64_t tile = t0; tile < t1; ++tile) {
    int seg;
    int64_t begin, end;
    if (!tile_range(g, segs, tile, seg, begin, end)) break;
    if (seg != segNow) {
      if (segNow >= 0) {
        __syncthreads();
        if (threadIdx.x < g.P && h[threadIdx.x]) atomicAdd(&hist[static_cast<int64_t>(segNow) * g.P + threadIdx.x], h[threadIdx.x]);
        if (threadIdx.x < kMaxP) h[threadIdx.x] = 0;
        __syncthreads();
      }
      segNow = seg;
    }
    const TileKeys tk = tile_keys(key, begin);
    const int rows = static_cast<int>(end - begin);
    constexpr int kRows = kTile / kPT;
    uint64_t k[kRows];
    if (rows == kTile) {  // full tile: every load in flight at once, no bounds checks
#pragma unroll
      for (int u = 0; u < kRows; ++u) k[u] = tk.p[threadIdx.x + u * kPT] - tk.sub;
    } else {
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        const int j = threadIdx.x + u * kPT;
        k[u] = j < rows ? tk.p[j] - tk.sub : 0;
      }
    }
#pragma unroll
    for (int u = 0; u < kRows; ++u) {  // block-uniform trip count: the ballots below see whole warps
      const bool live = static_cast<int>(threadIdx.x) + u * kPT < rows;
      const uint64_t mix = slice_mix(k[u]);
      if (live) atomicAdd(&h[(mix >> g.shift) & (g.P - 1)], 1u);
      if (hll) {
        // one key in eight feeds the sketch, through the strong hash (the convention of radix_partition.cu: x 8 on the host)
        const bool sampled = live && ((mix >> 20) & 7u) == 0;
        const unsigned m = __ballot_sync(0xffffffffu, sampled);
        if (sampled) hq[warp][qn + __popc(m & ((1u << lane) - 1u))] = k[u];
        qn += __popc(m);
        if (qn >= 32) {
          __syncwarp();
          sketch(hq[warp][qn - 32 + lane]);
          qn -= 32;
          __syncwarp();
        }
      }
    }
  }
  if (hll) {
    __syncwarp();
    if (lane < qn) sketch(hq[warp][lane]);
  }
  __syncthreads();
  if (segNow >= 0 && threadIdx.x < g.P && h[threadIdx.x]) atomicAdd(&hist[static_cast<int64_t>(segNow) * g.P + threadIdx.x], h[threadIdx.x]);
  if (hll)
    for (int i = threadIdx.x; i < (1 << kHllBits); i += kPT)
      if (regs[i]) atomicMax(&hll[i], regs[i]);
}

// level 1: partition starts, scatter cursors and the tile numbering of level 2
__global__ void part_scan1_kernel(const uint32_t* __restrict__ hist, int64_t* __restrict__ start, unsigned long long* __restrict__ cursor,
                                  int32_t* __restrict__ tile_start) {
  if (threadIdx.x != 0) return;
  int64_t run = 0;
  int32_t tiles = 0;
  for (int p = 0; p < kP1; ++p) {
    start[p] = run;
    cursor[p] = static_cast<unsigned long long>(run);
    tile_start[p] = tiles;
    run += hist[p];
    tiles += static_cast<int32_t>((hist[p] + kTile - 1) / kTile);
  }
  start[kP1] = run;
  tile_start[kP1] = tiles;
}
// level 2: slice starts inside every level-1 partition (block = partition)
__global__ void part_scan2_kernel(const uint32_t* __restrict__ hist, const int64_t* __restrict__ seg_start, int P, int64_t* __restrict__ slice_start,
                                  unsigned long long* __restrict__ cursor) {
  if (threadIdx.x != 0) return;
  const int s = blockIdx.x;
  int64_t run = seg_start[s];
  for (int d = 0; d < P; ++d) {
    slice_start[static_cast<int64_t>(s) * P + d] = run;
    cursor[static_cast<int64_t>(s) * P + d] = static_cast<unsigned long long>(run);
    run += hist[static_cast<int64_t>(s) * P + d];
  }
  if (s == gridDim.x - 1) slice_start[static_cast<int64_t>(gridDim.x) * P] = run;
}

// Staged scatter of one tile per iteration: (A) every thread loads its seven rows (keys, and payloads when
// they fit the register budget) in one burst, digits + tile histogram, (B) local offsets and the tile's
// reservation in every partition (one global atomic per non-empty digit), (C) keys and payloads from the
// registers into shared memory in partition order, (D) consecutive threads store consecutive rows.
// The input is read exactly once; its load latency is paid once per tile with all loads in flight. Full tiles
// (all but the last of a chunk / segment) run without bounds checks; with one payload column a staged row is one
// 16-byte shared-memory word (one STS.128 / LDS.128 instead of two 8-byte accesses each way).
template <int NCOLS, int MINB>
__global__ void __launch_bounds__(kPT, MINB) part_scatter_kernel(const __grid_constant__ PartIO io, const __grid_constant__ PartGeom g, int64_t ntiles,
                                                                               unsigned long long* __restrict__ cursor) {
  constexpr int kRows = kTile / kPT;             // rows per thread and tile
  constexpr bool kEarly = NCOLS <= 2;            // payloads loaded with the keys (three payload columns would spill)
  constexpr bool kPair = NCOLS == 1;             // staged rows are (key, payload) pairs
  static_assert(kRows * kPT == kTile && kRows <= 8, "a tile is a whole number of rows per thread; digits pack into one word");
  extern __shared__ __align__(16) uint8_t smem[];
  uint64_t* skeys = reinterpret_cast<uint64_t*>(smem);
  uint64_t* scols = skeys + kTile;                                           // [ncols][kTile]
  ulonglong2* spair = reinterpret_cast<ulonglong2*>(smem);                   // kPair: [kTile] (key, payload)
  uint8_t* sdig = reinterpret_cast<uint8_t*>(scols + static_cast<size_t>(NCOLS) * kTile);  // digit of every staged position
  __shared__ uint32_t cnt[kMaxP], lcur[kMaxP];
  __shared__ unsigned long long gdelta[kMaxP];   // global position of staged position p of digit d = gdelta[d] + p
  __shared__ uint32_t wtot[kMaxP / 32];
  __shared__ SegTable segs;
  load_seg_table(g, segs);
  const int tid = threadIdx.x;
  const int P = g.P;
  for (int64_t tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    int seg;
    int64_t begin, end;
    if (!tile_range(g, segs, tile, seg, begin, end)) break;
    const int rows = static_cast<int>(end - begin);
    const bool full = rows == kTile;
    if (tid < kMaxP) cnt[tid] = 0;
    // (A) the loads first: they fly while the counters are cleared
    const TileKeys tk = tile_keys(io.key, begin);
    uint64_t k[kRows], v[NCOLS > 0 ? NCOLS : 1][kRows];  // every index is a compile-time constant after unrolling: registers, no stack
    if (full) {
#pragma unroll
      for (int u = 0; u < kRows; ++u) k[u] = tk.p[tid + u * kPT] - tk.sub;
      if (kEarly) {
#pragma unroll
        for (int c = 0; c < NCOLS; ++c) {
          const uint64_t* cp = io.cols_in[c] + begin;
#pragma unroll
          for (int u = 0; u < kRows; ++u) v[c][u] = cp[tid + u * kPT];
        }
      }
    } else {
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        const int j = tid + u * kPT;
        k[u] = j < rows ? tk.p[j] - tk.sub : 0;
        if (kEarly) {
#pragma unroll
          for (int c = 0; c < NCOLS; ++c) v[c][u] = j < rows ? io.cols_in[c][begin + j] : 0;
        }
      }
    }
    // (An L2 prefetch of this block's next tile at this point was measured: no gain at level 1, 5 % slower at level 2.)
    __syncthreads();
    uint64_t digits = 0;  // one byte per row of this thread
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const uint32_t d = static_cast<uint32_t>((slice_mix(k[u]) >> g.shift) & (P - 1));
      digits |= static_cast<uint64_t>(d) << (8 * u);
      if (full || tid + u * kPT < rows) atomicAdd(&cnt[d], 1u);
    }
    __syncthreads();
    // (B) exclusive prefix of the digit counts (the first 256 threads: warp scans + warp totals)
    uint32_t mine = 0, incl = 0;
    if (tid < kMaxP) {
      mine = tid < P ? cnt[tid] : 0;
      incl = mine;
      const int lane = tid & 31;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
      }
      if (lane == 31) wtot[tid >> 5] = incl;
    }
    __syncthreads();
    if (tid < kMaxP) {
      uint32_t before = 0;
      for (int w = 0; w < (tid >> 5); ++w) before += wtot[w];
      const uint32_t lofs = before + incl - mine;
      lcur[tid] = lofs;
      if (mine) gdelta[tid] = atomicAdd(&cursor[static_cast<int64_t>(seg) * P + tid], static_cast<unsigned long long>(mine)) - lofs;
    }
    if (!kEarly) {
#pragma unroll
      for (int u = 0; u < kRows; ++u) {
        const int j = tid + u * kPT;
#pragma unroll
        for (int c = 0; c < NCOLS; ++c) v[c][u] = j < rows ? io.cols_in[c][begin + j] : 0;
      }
    }
    __syncthreads();
    // (C) placement
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      if (full || tid + u * kPT < rows) {
        const uint32_t d = static_cast<uint32_t>(digits >> (8 * u)) & 0xffu;
        const uint32_t p = atomicAdd(&lcur[d], 1u);
        sdig[p] = static_cast<uint8_t>(d);
        if (kPair) {
          spair[p] = make_ulonglong2(k[u], v[0][u]);
        } else {
          skeys[p] = k[u];
#pragma unroll
          for (int c = 0; c < NCOLS; ++c) scols[static_cast<size_t>(c) * kTile + p] = v[c][u];
        }
      }
    }
    __syncthreads();
    // (D)
#pragma unroll
    for (int u = 0; u < kRows; ++u) {
      const int p = tid + u * kPT;
      if (full || p < rows) {
        const unsigned long long out = gdelta[sdig[p]] + static_cast<unsigned>(p);
        if (kPair) {
          const ulonglong2 r = spair[p];
          io.keys_out[out] = r.x;
          io.cols_out[0][out] = r.y;
        } else {
          io.keys_out[out] = skeys[p];
#pragma unroll
          for (int c = 0; c < NCOLS; ++c) io.cols_out[c][out] = scols[static_cast<size_t>(c) * kTile + p];
        }
      }
    }
    __syncthreads();
  }
}

// ---- aggregate -------------------------------------------------------------------------------------
struct AggIO {
  const uint64_t* keys;
  const uint64_t* cols[kCols];
  const int64_t* slice_start;  // [nslices + 1]
  int nslices;
  vb2_slice_op ops[kMaxOps];
  int nops;
  uint64_t row_init[VB2_MAX_ROW_WORDS];
  int row_words;
  uint64_t* rows_out;
  int64_t rows_capacity;
  unsigned long long* num_groups;  // groups written
  unsigned long long* reserved;    // output rows handed out to the blocks (chunks; >= groups)
  int64_t chunk_rows;
  int32_t* error_flag;     // 1 = SUM(BIGINT) overflow, 100 = rows_out full
  int32_t* overflow;       // slices whose distinct keys did not fit the shared-memory table
  int C;                   // slots per slice table (power of two)
};

// ---- shared-memory access by 32-bit shared address --------------------------------------------------
// Through generic pointers the compiler rebuilds the CTA's shared window base (S2R SR_CgaCtaId + LEA) in front
// of shared accesses of the probe loop; with the 32-bit address held in a register the loop is LDS / ATOMS only.
__device__ __forceinline__ uint64_t lds_volatile_u64(uint32_t addr) {
  uint64_t v;
  asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t atoms_cas_u64(uint32_t addr, uint64_t expect, uint64_t desired) {
  uint64_t old;
  asm volatile("atom.shared.cas.b64 %0, [%1], %2, %3;" : "=l"(old) : "r"(addr), "l"(expect), "l"(desired) : "memory");
  return old;
}
__device__ __forceinline__ uint32_t atoms_add_u32(uint32_t addr, uint32_t x) {
  uint32_t old;
  asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(addr), "r"(x) : "memory");
  return old;
}
__device__ __forceinline__ void reds_add_u32(uint32_t addr, uint32_t x) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(addr), "r"(x) : "memory"); }
__device__ __forceinline__ uint64_t atoms_add_u64(uint32_t addr, uint64_t x) {
  uint64_t old;
  asm volatile("atom.shared.add.u64 %0, [%1], %2;" : "=l"(old) : "r"(addr), "l"(x) : "memory");
  return old;
}
__device__ __forceinline__ void reds_add_f64(uint32_t addr, double x) { asm volatile("red.shared.add.f64 [%0], %1;" ::"r"(addr), "d"(x) : "memory"); }
__device__ __forceinline__ void reds_minmax_s64(uint32_t addr, int64_t x, bool is_min) {
  if (is_min) asm volatile("red.shared.min.s64 [%0], %1;" ::"r"(addr), "l"(x) : "memory");
  else asm volatile("red.shared.max.s64 [%0], %1;" ::"r"(addr), "l"(x) : "memory");
}
__device__ __forceinline__ void smem_minmax_f64(uint32_t addr, double v, bool is_min) {
  uint64_t old = lds_volatile_u64(addr);
  for (;;) {
    const double cur = __longlong_as_double(static_cast<long long>(old));
    const bool better = is_min ? lt_f64(v, cur) : gt_f64(v, cur);
    if (!better) return;
    const uint64_t seen = atoms_cas_u64(addr, old, static_cast<uint64_t>(__double_as_longlong(v)));
    if (seen == old) return;
    old = seen;
  }
}

// One accumulator update in shared memory (acc = 32-bit shared address of the word). 64-bit shared-memory
// integer adds are compare-and-swap loops in SASS (LDS + ATOMS.CAST.SPIN); 32-bit ones are native, so integer sums
// are carried as two halves: the low add returns the old half, the carry (and the sign extension) reach the high
// half only when non-zero — for small values almost never. A slice holds < 2^32 rows of |x| < 2^31: that 64-bit sum
// cannot overflow.
__device__ __forceinline__ void slice_update(int kind, uint32_t acc, uint64_t raw, int32_t* error_flag) {
  switch (kind) {
    case VB2_AGG_SUM_F64: reds_add_f64(acc, __longlong_as_double(static_cast<long long>(raw))); break;
    case VB2_AGG_SUM_I64: case VB2_AGG_COUNT_MERGE: {
      const int64_t x = static_cast<int64_t>(raw);
      if (x == static_cast<int32_t>(x)) {
        const uint32_t xl = static_cast<uint32_t>(x);
        const uint32_t old = atoms_add_u32(acc, xl);
        const uint32_t up = static_cast<uint32_t>(x >> 32) + (static_cast<uint32_t>(old + xl) < old ? 1u : 0u);
        if (up) reds_add_u32(acc + 4, up);
      } else {
        const int64_t old = static_cast<int64_t>(atoms_add_u64(acc, static_cast<uint64_t>(x)));
        int64_t r;
        if (add_overflow_i64(old, x, &r)) atomicCAS(error_flag, 0, 1);
      }
      break;
    }
    case VB2_AGG_COUNT: reds_add_u32(acc, 1u); break;  // < 2^32 rows per slice: the low half suffices
    case VB2_AGG_MIN_F64: smem_minmax_f64(acc, __longlong_as_double(static_cast<long long>(raw)), true); break;
    case VB2_AGG_MAX_F64: smem_minmax_f64(acc, __longlong_as_double(static_cast<long long>(raw)), false); break;
    case VB2_AGG_MIN_I64: reds_minmax_s64(acc, static_cast<int64_t>(raw), true); break;
    case VB2_AGG_MAX_I64: reds_minmax_s64(acc, static_cast<int64_t>(raw), false); break;
    default: break;
  }
}

// Payload column of op o in the instantiations that fix their kinds: the o-th op that reads a column reads column o'
// = number of column-reading ops before it (the host checks the op list has that shape, else the generic build runs).
__host__ __device__ constexpr int fixed_col(uint32_t kinds, int o) {
  int c = 0;
  for (int i = 0; i < o; ++i)
    if (((kinds >> (4 * i)) & 15u) != VB2_AGG_COUNT) ++c;
  return c;
}

// NOPS accumulator words per slot (compile time: the op descriptors live in registers, the loops unroll).
// KINDS: four bits per op holding its vb2_agg_kind when the instantiation fixes it (the update switch and the
// column select fold away), 0 = read from the descriptor at run time. The generic build carries every case of
// every op four times over; the common op lists get their own straight-line builds.
// Output rows are reserved from the global cursor in chunks: a block asks for a new chunk only when the
// groups of its next slice do not fit the rest of its current one, so the round trip of a global atomic is
// paid once per few dozen slices; the unused tail of a chunk stays EMPTY rows of the table-shaped output.
template <int NOPS, uint32_t KINDS>
__global__ void __launch_bounds__(kAggThreads) slice_aggregate_kernel(const __grid_constant__ AggIO a) {
  extern __shared__ __align__(16) uint8_t smem[];
  uint64_t* skey = reinterpret_cast<uint64_t*>(smem);  // [C]
  uint64_t* sacc = skey + a.C;                          // [NOPS][C]
  __shared__ int s_overflow;
  __shared__ unsigned int s_count, s_cursor;
  __shared__ unsigned long long s_base, s_chunk_pos, s_chunk_left, s_groups;
  const int tid = threadIdx.x;
  const int lane = tid & 31;
  const int C = a.C;
  const uint32_t cmask = static_cast<uint32_t>(C - 1);
  uint32_t skey_s = static_cast<uint32_t>(__cvta_generic_to_shared(skey));
  asm volatile("mov.u32 %0, %0;" : "+r"(skey_s));  // opaque: kept in a register instead of being rebuilt from SR_CgaCtaId at every use
  const uint32_t sacc_s = skey_s + static_cast<uint32_t>(C) * 8u;
  int kind[NOPS], col[NOPS], word[NOPS];
  uint64_t init[NOPS];
#pragma unroll
  for (int o = 0; o < NOPS; ++o) {
    kind[o] = a.ops[o].kind;
    col[o] = a.ops[o].col;
    word[o] = a.ops[o].word;
    init[o] = a.row_init[a.ops[o].word];
  }
  if (tid == 0) { s_chunk_pos = 0; s_chunk_left = 0; s_groups = 0; }
  for (int s = blockIdx.x; s < a.nslices; s += gridDim.x) {
    const int64_t begin = a.slice_start[s], end = a.slice_start[s + 1];
    if (begin == end) continue;  // uniform across the block
    // The first rows of the slice are requested before the table is cleared, and the rows of iteration i + 1 before
    // iteration i is processed: the global-load latency (the largest single stall of the one-buffer version,
    // paid by every warp at the same time behind the per-slice barriers) overlaps shared-memory work.
    uint64_t nk[4], nv[kCols][4];
    auto request = [&](int64_t i0) {
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + static_cast<int64_t>(u) * kAggThreads;
        if (i < end) {
          nk[u] = a.keys[i];
#pragma unroll
          for (int c = 0; c < kCols; ++c)
            if (a.cols[c]) nv[c][u] = a.cols[c][i];
        }
      }
    };
    request(begin + tid);
    for (int i = tid; i < C; i += kAggThreads) {
      skey[i] = VB2_EMPTY_KEY;
#pragma unroll
      for (int o = 0; o < NOPS; ++o) sacc[o * C + i] = init[o];
    }
    if (tid == 0) { s_overflow = 0; s_count = 0; s_cursor = 0; }
    __syncthreads();
    int fresh = 0;  // groups this thread inserted
    // warp-uniform trip count (the probe loop below votes): every lane of a warp walks the same number of iterations
    const int64_t warp0 = begin + (tid & ~31);
    for (int64_t w0 = warp0; w0 < end; w0 += 4 * kAggThreads) {
      const int64_t i0 = w0 + lane;
      uint64_t k[4], v[kCols][4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        k[u] = nk[u];
#pragma unroll
        for (int c = 0; c < kCols; ++c) v[c][u] = nv[c][u];
      }
      request(i0 + 4 * kAggThreads);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int64_t i = i0 + static_cast<int64_t>(u) * kAggThreads;
        const bool live = i < end;
        const uint64_t key = k[u];
        uint32_t slot = static_cast<uint32_t>(slice_mix(key) >> 34) & cmask;  // bits below the ones that chose the slice
        // Convergent probe: the warp loops until every live lane has its slot, then all lanes update together
        // (with the exit inside the loop the compiler duplicated the update into it and ran it once per probe round,
        // half the lanes idle).
        bool done = !live;
        int probes = 0;
        while (!__all_sync(0xffffffffu, done)) {
          if (!done) {
            uint64_t cur = lds_volatile_u64(skey_s + slot * 8u);
            if (cur == VB2_EMPTY_KEY) {
              cur = atoms_cas_u64(skey_s + slot * 8u, VB2_EMPTY_KEY, key);
              if (cur == VB2_EMPTY_KEY) { ++fresh; cur = key; }  // a new group of this slice
            }
            if (cur == key) done = true;
            else if (++probes >= C) { s_overflow = 1; done = true; slot = 0xffffffffu; }
            else slot = (slot + 1) & cmask;
          }
        }
        if (!live || slot == 0xffffffffu) continue;
#pragma unroll
        for (int o = 0; o < NOPS; ++o) {
          const int fixed = static_cast<int>((KINDS >> (4 * o)) & 15u);  // constants once the op loop is unrolled
          const int fcol = fixed_col(KINDS, o);
          const int cl = KINDS ? fcol : col[o];
          const uint64_t raw = cl == 0 ? v[0][u] : (cl == 1 ? v[1][u] : (cl == 2 ? v[2][u] : 0));  // constant indices: registers
          slice_update(fixed ? fixed : kind[o], sacc_s + (static_cast<uint32_t>(o) * C + slot) * 8u, raw, a.error_flag);
        }
      }
    }
    fresh = warp_sum(fresh);
    if (lane == 0 && fresh) atomicAdd(&s_count, static_cast<unsigned>(fresh));
    __syncthreads();
    if (tid == 0) {
      const unsigned long long n = s_count;
      if (n > s_chunk_left) {
        const unsigned long long grab = n > static_cast<unsigned long long>(a.chunk_rows) ? n : static_cast<unsigned long long>(a.chunk_rows);
        s_chunk_pos = atomicAdd(a.reserved, grab);
        s_chunk_left = grab;
      }
      s_base = s_chunk_pos;
      s_chunk_pos += n;
      s_chunk_left -= n;
      s_groups += n;
      if (s_overflow) atomicAdd(a.overflow, 1);
      if (static_cast<int64_t>(s_base + n) > a.rows_capacity) atomicCAS(a.error_flag, 0, 100);
    }
    __syncthreads();
    if (static_cast<int64_t>(s_base + s_count) <= a.rows_capacity) {
      for (int i0 = 0; i0 < C; i0 += kAggThreads) {  // block-uniform trip count: whole warps reach the ballot
        const int i = i0 + tid;
        const uint64_t key = i < C ? skey[i] : VB2_EMPTY_KEY;
        const bool live = key != VB2_EMPTY_KEY;
        // one shared-memory atomic per warp: the occupied slots of a warp take consecutive output rows
        const unsigned m = __ballot_sync(0xffffffffu, live);
        unsigned first = 0;
        if (lane == 0 && m) first = atomicAdd(&s_cursor, static_cast<unsigned>(__popc(m)));
        first = __shfl_sync(0xffffffffu, first, 0);
        if (!live) continue;
        uint64_t* row = a.rows_out + (s_base + first + __popc(m & ((1u << lane) - 1u))) * static_cast<unsigned long long>(a.row_words);
        row[0] = key;
        for (int w = 1; w < a.row_words; ++w) row[w] = a.row_init[w];
#pragma unroll
        for (int o = 0; o < NOPS; ++o) row[word[o]] = sacc[o * C + i];
      }
    }
    __syncthreads();
  }
  if (tid == 0 && s_groups) atomicAdd(a.num_groups, s_groups);
}

inline size_t align256(size_t v) { return (v + 255) / 256 * 256; }

// workspace carving shared by the two entry points
struct Workspace {
  uint64_t *keysA, *keysB;
  uint64_t *colsA[kCols], *colsB[kCols];
  uint32_t* hist1;
  int64_t* start1;
  unsigned long long* cursor1;
  int32_t* tile_start2;
  uint32_t* hist2;
  int64_t* slice_start;
  unsigned long long* cursor2;
  int32_t* hll;
  size_t bytes;
};
Workspace carve(void* base, int64_t n, int ncols) {
  Workspace w{};
  uint8_t* p = static_cast<uint8_t*>(base);
  size_t off = 0;
  auto take = [&](size_t b) {
    uint8_t* r = p ? p + off : nullptr;
    off += align256(b);
    return r;
  };
  const size_t col = static_cast<size_t>(n) * 8;
  w.keysA = reinterpret_cast<uint64_t*>(take(col));
  w.keysB = reinterpret_cast<uint64_t*>(take(col));
  for (int c = 0; c < ncols; ++c) {
    w.colsA[c] = reinterpret_cast<uint64_t*>(take(col));
    w.colsB[c] = reinterpret_cast<uint64_t*>(take(col));
  }
  w.hist1 = reinterpret_cast<uint32_t*>(take(kP1 * 4));
  w.start1 = reinterpret_cast<int64_t*>(take((kP1 + 1) * 8));
  w.cursor1 = reinterpret_cast<unsigned long long*>(take(kP1 * 8));
  w.tile_start2 = reinterpret_cast<int32_t*>(take((kP1 + 1) * 4));
  w.hist2 = reinterpret_cast<uint32_t*>(take(static_cast<size_t>(kP1) * kMaxP * 4));
  w.slice_start = reinterpret_cast<int64_t*>(take((static_cast<size_t>(kP1) * kMaxP + 1) * 8));
  w.cursor2 = reinterpret_cast<unsigned long long*>(take(static_cast<size_t>(kP1) * kMaxP * 8));
  w.hll = reinterpret_cast<int32_t*>(take((1 << kHllBits) * 4));
  w.bytes = off;
  return w;
}

// rows a block reserves from the output cursor at a time: a few chunks per block over the whole run
int64_t output_chunk_rows(int64_t distinct_estimate) {
  const int64_t per = distinct_estimate / (static_cast<int64_t>(device_sm_count()) * 2 * 4);
  return per < 256 ? 256 : (per > 16384 ? 16384 : per);
}
size_t scatter_smem(int ncols) { return static_cast<size_t>(kTile) * 8 * (1 + ncols) + kTile; }
// CTAs per SM of the scatter: two (spill-free at 56-64 registers; measured 2.4 ms faster over 1 B rows than three
// CTAs under a 42-register cap that spills a few words per tile). VB2_SLICE_SCATTER_CTAS=3 selects the three-CTA
// build of the narrow kernels (measurement aid).
int scatter_ctas(int ncols) {
  static const int forced = [] { const char* e = std::getenv("VB2_SLICE_SCATTER_CTAS"); return e ? std::atoi(e) : 0; }();
  if (ncols > 1) return 2;
  return forced == 3 ? 3 : 2;
}
int configure_scatter() {
  static bool configured = false;
  if (configured) return VB2_OK;
  VB2_CUDA_OK(cudaFuncSetAttribute(part_scatter_kernel<0, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scatter_smem(0))));
  VB2_CUDA_OK(cudaFuncSetAttribute(part_scatter_kernel<1, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scatter_smem(1))));
  VB2_CUDA_OK(cudaFuncSetAttribute(part_scatter_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scatter_smem(0))));
  VB2_CUDA_OK(cudaFuncSetAttribute(part_scatter_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scatter_smem(1))));
  VB2_CUDA_OK(cudaFuncSetAttribute(part_scatter_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scatter_smem(2))));
  VB2_CUDA_OK(cudaFuncSetAttribute(part_scatter_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(scatter_smem(3))));
  configured = true;
  return VB2_OK;
}
void launch_scatter(const PartIO& io, const PartGeom& g, int64_t ntiles, unsigned long long* cursor, unsigned grid, cudaStream_t st) {
  const size_t smem = scatter_smem(io.ncols);
  const bool three = scatter_ctas(io.ncols) == 3;
  switch (io.ncols) {
    case 0:
      if (three) part_scatter_kernel<0, 3><<<counted(grid), kPT, smem, st>>>(io, g, ntiles, cursor);
      else part_scatter_kernel<0, 2><<<counted(grid), kPT, smem, st>>>(io, g, ntiles, cursor);
      break;
    case 1:
      if (three) part_scatter_kernel<1, 3><<<counted(grid), kPT, smem, st>>>(io, g, ntiles, cursor);
      else part_scatter_kernel<1, 2><<<counted(grid), kPT, smem, st>>>(io, g, ntiles, cursor);
      break;
    case 2: part_scatter_kernel<2, 2><<<counted(grid), kPT, smem, st>>>(io, g, ntiles, cursor); break;
    default: part_scatter_kernel<3, 2><<<counted(grid), kPT, smem, st>>>(io, g, ntiles, cursor);
  }
}

}  // namespace
}  // namespace vb2

using namespace vb2;

extern "C" {

int32_t vb2k_slice_agg_hll_registers(void) { return 1 << kHllBits; }

size_t vb2k_slice_agg_workspace(int64_t total_rows, int32_t ncols) { return carve(nullptr, total_rows, ncols).bytes; }

int vb2k_slice_agg_partition(const vb2_slice_chunk* chunks, int32_t nchunks, int32_t ncols, int64_t total_rows, void* workspace, size_t workspace_bytes,
                             int32_t* hll_host, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (ncols < 0 || ncols > kCols) return fail_msg(VB2_ERR_UNSUPPORTED, "slice aggregation: at most 3 payload columns");
  if (total_rows >= (1ll << 32)) return fail_msg(VB2_ERR_UNSUPPORTED, "slice aggregation: above 2^32 rows");
  Workspace w = carve(workspace, total_rows, ncols);
  if (workspace_bytes < w.bytes) return fail_msg(VB2_ERR_INVALID, "slice aggregation: workspace too small");
  int64_t sum = 0;
  for (int i = 0; i < nchunks; ++i) sum += chunks[i].rows;
  if (sum != total_rows) return fail_msg(VB2_ERR_INVALID, "slice aggregation: chunk rows do not add up");
  VB2_CUDA_OK(cudaMemsetAsync(w.hist1, 0, kP1 * 4, st));
  VB2_CUDA_OK(cudaMemsetAsync(w.hll, 0, (1 << kHllBits) * 4, st));
  if (int rc = configure_scatter()) return rc;
  const int sms = device_sm_count();
  auto geom_of = [&](int64_t n) {
    PartGeom g{};
    g.n = n;
    g.nseg = 1;
    g.P = kP1;
    g.shift = 56;
    return g;
  };
  for (int i = 0; i < nchunks; ++i) {
    if (chunks[i].rows == 0) continue;
    const SliceKey key{chunks[i].norm_keys, chunks[i].raw_keys, chunks[i].key_min};
    const int64_t ntiles = (chunks[i].rows + kTile - 1) / kTile;
    const int64_t cap = static_cast<int64_t>(sms) * 4;
    part_hist_kernel<<<counted(static_cast<unsigned>(ntiles < cap ? ntiles : cap)), kPT, 0, st>>>(key, geom_of(chunks[i].rows), ntiles, w.hist1, w.hll);
  }
  part_scan1_kernel<<<counted(1u), 32, 0, st>>>(w.hist1, w.start1, w.cursor1, w.tile_start2);
  for (int i = 0; i < nchunks; ++i) {
    if (chunks[i].rows == 0) continue;
    PartIO io{};
    io.key = SliceKey{chunks[i].norm_keys, chunks[i].raw_keys, chunks[i].key_min};
    io.ncols = ncols;
    io.keys_out = w.keysA;
    for (int c = 0; c < ncols; ++c) {
      io.cols_in[c] = static_cast<const uint64_t*>(chunks[i].cols[c]);
      io.cols_out[c] = w.colsA[c];
    }
    const int64_t ntiles = (chunks[i].rows + kTile - 1) / kTile;
    const int64_t cap = static_cast<int64_t>(sms) * scatter_ctas(ncols);
    launch_scatter(io, geom_of(chunks[i].rows), ntiles, w.cursor1, static_cast<unsigned>(ntiles < cap ? ntiles : cap), st);
  }
  VB2_CUDA_OK(cudaGetLastError());
  if (hll_host) {
    VB2_CUDA_OK(cudaMemcpyAsync(hll_host, w.hll, (1 << kHllBits) * 4, cudaMemcpyDeviceToHost, st));
    VB2_CUDA_OK(cudaStreamSynchronize(st));
  }
  return VB2_OK;
}

int64_t vb2k_slice_agg_output_rows(int64_t distinct_estimate) {
  return distinct_estimate + static_cast<int64_t>(device_sm_count()) * 2 * output_chunk_rows(distinct_estimate) + 1024;
}

int vb2k_slice_agg_finish(int64_t total_rows, int32_t ncols, int64_t distinct_estimate, const vb2_slice_op* ops, int32_t nops, int32_t row_words,
                          const uint64_t* row_init, uint64_t* rows_out, int64_t rows_capacity, int64_t* num_groups, int64_t* reserved_rows,
                          int32_t* error_flag, int32_t* overflow_slices, void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (nops < 1 || nops > kMaxOps || row_words < 2 || row_words > VB2_MAX_ROW_WORDS) return fail_msg(VB2_ERR_UNSUPPORTED, "slice aggregation: 1 to 8 accumulator words");
  Workspace w = carve(workspace, total_rows, ncols);
  if (workspace_bytes < w.bytes) return fail_msg(VB2_ERR_INVALID, "slice aggregation: workspace too small");
  // Slot count of a slice table (keys + one word per op): the largest power of two within ~100 KB so that
  // two CTAs share an SM; one size up (one CTA per SM) when 65536 slices would otherwise load their
  // tables above one half. Slices: expected distinct keys per slice <= C / 2.
  int C = 8192;
  while (C > 256 && static_cast<size_t>(C) * 8 * (1 + nops) > 100 * 1024) C >>= 1;
  auto slices_for = [&](int c) { return (distinct_estimate * 2 + c - 1) / c; };
  if (slices_for(C) > static_cast<int64_t>(kP1) * kMaxP && static_cast<size_t>(C) * 2 * 8 * (1 + nops) <= 200 * 1024) C <<= 1;
  const int64_t want = slices_for(C);
  int P2 = 1;
  while (static_cast<int64_t>(kP1) * P2 < want && P2 < kMaxP) P2 <<= 1;
  if (static_cast<int64_t>(kP1) * P2 < want) return fail_msg(VB2_ERR_UNSUPPORTED, "slice aggregation: too many distinct keys for 65536 shared-memory slices");
  const int sms = device_sm_count();
  AggIO a{};
  if (P2 > 1) {
    PartGeom g{};
    g.n = total_rows;
    g.seg_start = w.start1;
    g.tile_start = w.tile_start2;
    g.nseg = kP1;
    g.P = P2;
    g.shift = 56;
    int bits = 0;
    while ((1 << bits) < P2) ++bits;
    g.shift = 56 - bits;
    const int64_t ntiles = total_rows / kTile + kP1 + 1;  // upper bound: every partition rounds its last tile up
    VB2_CUDA_OK(cudaMemsetAsync(w.hist2, 0, static_cast<size_t>(kP1) * P2 * 4, st));
    const SliceKey key{w.keysA, nullptr, 0};
    const int64_t hcap = static_cast<int64_t>(sms) * 4;
    part_hist_kernel<<<counted(static_cast<unsigned>(ntiles < hcap ? ntiles : hcap)), kPT, 0, st>>>(key, g, ntiles, w.hist2, nullptr);
    part_scan2_kernel<<<counted(static_cast<unsigned>(kP1)), 32, 0, st>>>(w.hist2, w.start1, P2, w.slice_start, w.cursor2);
    PartIO io{};
    io.key = key;
    io.ncols = ncols;
    io.keys_out = w.keysB;
    for (int c = 0; c < ncols; ++c) {
      io.cols_in[c] = w.colsA[c];
      io.cols_out[c] = w.colsB[c];
    }
    const int64_t scap = static_cast<int64_t>(sms) * scatter_ctas(ncols);
    launch_scatter(io, g, ntiles, w.cursor2, static_cast<unsigned>(ntiles < scap ? ntiles : scap), st);
    a.keys = w.keysB;
    for (int c = 0; c < ncols; ++c) a.cols[c] = w.colsB[c];
    a.slice_start = w.slice_start;
    a.nslices = kP1 * P2;
  } else {
    a.keys = w.keysA;
    for (int c = 0; c < ncols; ++c) a.cols[c] = w.colsA[c];
    a.slice_start = w.start1;
    a.nslices = kP1;
  }
  for (int o = 0; o < nops; ++o) {
    if (ops[o].col >= ncols || ops[o].word < 1 || ops[o].word >= row_words) return fail_msg(VB2_ERR_INVALID, "slice aggregation: bad accumulator op");
    a.ops[o] = ops[o];
  }
  a.nops = nops;
  for (int i = 0; i < row_words; ++i) a.row_init[i] = row_init[i];
  a.row_words = row_words;
  a.rows_out = rows_out;
  a.rows_capacity = rows_capacity;
  a.num_groups = reinterpret_cast<unsigned long long*>(num_groups);
  a.reserved = reinterpret_cast<unsigned long long*>(reserved_rows);
  a.chunk_rows = output_chunk_rows(distinct_estimate < total_rows ? distinct_estimate : total_rows);  // the caller sizes rows_out from the same bound
  a.error_flag = error_flag;
  a.overflow = overflow_slices;
  a.C = C;
  const size_t smem = static_cast<size_t>(C) * 8 * (1 + nops);
  const int64_t acap = static_cast<int64_t>(sms) * 2;
  const unsigned grid = static_cast<unsigned>(a.nslices < acap ? a.nslices : acap);
  // packed kinds of this op list (<= 8 ops x 4 bits)
  uint32_t packed = 0;
  for (int o = 0; o < nops; ++o) packed |= static_cast<uint32_t>(ops[o].kind & 15) << (4 * o);
  static const bool genericOnly = [] { const char* e = std::getenv("VB2_SLICE_GENERIC_AGG"); return e && e[0] == '1'; }();
  // the kind-specialised builds read op o's input from column fixed_col(kinds, o): only op lists of that shape take them
  bool columnsInOrder = true;
  for (int o = 0; o < nops; ++o)
    if (ops[o].kind != VB2_AGG_COUNT && ops[o].col != fixed_col(packed, o)) columnsInOrder = false;
#define VB2_SLICE_LAUNCH(N, K)                                                                                                       \
  {                                                                                                                                 \
    static size_t configured = 0;                                                                                                   \
    if (smem > configured) {                                                                                                        \
      VB2_CUDA_OK(cudaFuncSetAttribute(slice_aggregate_kernel<N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))); \
      configured = smem;                                                                                                            \
    }                                                                                                                               \
    slice_aggregate_kernel<N, K><<<counted(grid), kAggThreads, smem, st>>>(a);                                                      \
  }
#define VB2_SLICE_FIXED(N, K) \
  if (!launched && !genericOnly && columnsInOrder && nops == N && packed == K) { VB2_SLICE_LAUNCH(N, K) launched = true; }
#define VB2_SLICE_AGG(N) \
  case N: VB2_SLICE_LAUNCH(N, 0u) break;
  bool launched = false;
  // kinds: 1 SUM_F64, 2 SUM_I64, 3 COUNT, 4 / 5 MIN / MAX F64, 6 / 7 MIN / MAX I64, 8 COUNT_MERGE (op 0 in the low nibble)
  VB2_SLICE_FIXED(1, 0x1u) VB2_SLICE_FIXED(1, 0x2u) VB2_SLICE_FIXED(1, 0x3u) VB2_SLICE_FIXED(1, 0x8u)
  VB2_SLICE_FIXED(2, 0x32u) VB2_SLICE_FIXED(2, 0x31u) VB2_SLICE_FIXED(2, 0x82u) VB2_SLICE_FIXED(2, 0x81u)
  VB2_SLICE_FIXED(2, 0x11u) VB2_SLICE_FIXED(2, 0x22u) VB2_SLICE_FIXED(2, 0x54u) VB2_SLICE_FIXED(2, 0x76u)
  VB2_SLICE_FIXED(3, 0x332u) VB2_SLICE_FIXED(3, 0x331u) VB2_SLICE_FIXED(3, 0x882u) VB2_SLICE_FIXED(3, 0x881u)
  if (!launched)
    switch (nops) {
      VB2_SLICE_AGG(1) VB2_SLICE_AGG(2) VB2_SLICE_AGG(3) VB2_SLICE_AGG(4) VB2_SLICE_AGG(5) VB2_SLICE_AGG(6) VB2_SLICE_AGG(7) VB2_SLICE_AGG(8)
    }
#undef VB2_SLICE_AGG
#undef VB2_SLICE_FIXED
#undef VB2_SLICE_LAUNCH
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

}  // extern "C"
