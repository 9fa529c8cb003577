// Fused scan -> filter -> [join probe] -> project -> array-mode aggregate, specialised ahead of time.
//
// The expression DAG of a pipeline is a C++ expression template. The same template prints the
// canonical text used as the registry key (sig()), so the code that runs and the key the
// planner matches against cannot drift apart. Arithmetic uses the explicit round-to-nearest
// intrinsics (__dmul_rn ...) so that no FMA contraction happens: the reference performs one
// IEEE-754 rounding per function call (functions/prestosql/Arithmetic.h:52-141).
//
// Data movement (HBM-bound; no tensor cores on this path):
//   * every input column is read exactly once with 128-bit (f64/i64 pair) or 64-bit (i32 pair)
//     non-allocating loads; consecutive lanes take consecutive row pairs, so each warp-level load
//     is one fully used 512 B / 256 B span;
//   * kUnroll independent pairs per thread are in flight before the first use;
//   * group accumulators live in registers (predicated adds, <= 16 groups), are reduced with
//     shuffles, and each block writes one partial; a second tiny kernel folds the partials in
//     block order -> results are run-to-run deterministic;
//   * persistent grid: blocks = SMs x resident blocks per SM.
#pragma once
#ifndef __CUDACC_RTC__
#include <string>
#endif

#include "common.cuh"

// The whole header also compiles under NVRTC (expression-template pipelines are instantiated at
// run time for plan shapes without an ahead-of-time specialisation, fused_jit.cu): nothing here
// may depend on the host standard library except the signature printers, which NVRTC skips.
#ifdef __CUDACC_RTC__
#define VB2_SIG(...)
#else
#define VB2_SIG(...) __VA_ARGS__
#endif

namespace vb2 {
namespace fx {

constexpr int kMaxCols = VB2_FUSED_MAX_COLS;

// Registers holding one row pair of every referenced column.
struct PairRegs {
  double f[kMaxCols][2];
  int32_t i[kMaxCols][2];
  int64_t l[kMaxCols][2];
  bool join_flag[2];
};

struct Consts {
  double pf[VB2_FUSED_MAX_PARAMS];
  int64_t pl[VB2_FUSED_MAX_PARAMS];
  int32_t pi[VB2_FUSED_MAX_PARAMS];
};

template <class... Ts>
struct TypeList {
  static constexpr int size = sizeof...(Ts);
};
template <class A, class B> struct IsSame { static constexpr bool value = false; };
template <class A> struct IsSame<A, A> { static constexpr bool value = true; };
template <class A, class B> constexpr bool is_same_v = IsSame<A, B>::value;
template <int I, class T, class... Ts> struct TypeAt { using type = typename TypeAt<I - 1, Ts...>::type; };
template <class T, class... Ts> struct TypeAt<0, T, Ts...> { using type = T; };

// ---- leaves ---------------------------------------------------------------------------------
template <int C>
struct ColF {
  using T = double;
  static constexpr uint32_t fmask = 1u << C, imask = 0, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ double eval(const PairRegs& r, const Consts&, int s) { return r.f[C][s]; }
  VB2_SIG(static std::string sig() { return "f" + std::to_string(C); })
};
template <int C>
struct ColI {
  using T = int32_t;
  static constexpr uint32_t fmask = 0, imask = 1u << C, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ int32_t eval(const PairRegs& r, const Consts&, int s) { return r.i[C][s]; }
  VB2_SIG(static std::string sig() { return "i" + std::to_string(C); })
};
template <int C>
struct ColL {
  using T = int64_t;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 1u << C;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ int64_t eval(const PairRegs& r, const Consts&, int s) { return r.l[C][s]; }
  VB2_SIG(static std::string sig() { return "l" + std::to_string(C); })
};
template <int K>
struct PF {
  using T = double;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ double eval(const PairRegs&, const Consts& c, int) { return c.pf[K]; }
  VB2_SIG(static std::string sig() { return "pf" + std::to_string(K); })
};
template <int K>
struct PI {
  using T = int32_t;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ int32_t eval(const PairRegs&, const Consts& c, int) { return c.pi[K]; }
  VB2_SIG(static std::string sig() { return "pi" + std::to_string(K); })
};
template <int K>
struct PL {
  using T = int64_t;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ int64_t eval(const PairRegs&, const Consts& c, int) { return c.pl[K]; }
  VB2_SIG(static std::string sig() { return "pl" + std::to_string(K); })
};
struct True {
  using T = bool;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ bool eval(const PairRegs&, const Consts&, int) { return true; }
  VB2_SIG(static std::string sig() { return "true"; })
};
// Build-side predicate of the matched build row (e.g. p_type LIKE 'PROMO%'), evaluated once per
// dictionary entry on the build side and looked up here.
struct JoinFlag {
  using T = bool;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 0;
  static constexpr bool uses_join = true;
  __device__ static __forceinline__ bool eval(const PairRegs& r, const Consts&, int s) { return r.join_flag[s]; }
  VB2_SIG(static std::string sig() { return "joinflag"; })
};

// ---- operators ------------------------------------------------------------------------------
#define VB2_FX_BINARY(Name, text, expr)                                                          \
  template <class A, class B>                                                                    \
  struct Name {                                                                                  \
    static_assert(is_same_v<typename A::T, double> && is_same_v<typename B::T, double>); \
    using T = double;                                                                            \
    static constexpr uint32_t fmask = A::fmask | B::fmask, imask = A::imask | B::imask,          \
                              lmask = A::lmask | B::lmask;                                       \
    static constexpr bool uses_join = A::uses_join || B::uses_join;                              \
    __device__ static __forceinline__ double eval(const PairRegs& r, const Consts& c, int s) {   \
      const double a = A::eval(r, c, s), b = B::eval(r, c, s);                                   \
      return expr;                                                                               \
    }                                                                                            \
    VB2_SIG(static std::string sig() { return std::string(text "(") + A::sig() + "," + B::sig() + ")"; }) \
  };
VB2_FX_BINARY(Plus, "plus", __dadd_rn(a, b))
VB2_FX_BINARY(Minus, "minus", __dsub_rn(a, b))
VB2_FX_BINARY(Multiply, "multiply", __dmul_rn(a, b))
VB2_FX_BINARY(Divide, "divide", __ddiv_rn(a, b))
#undef VB2_FX_BINARY

template <int Op, class A, class B>
struct Compare {
  static_assert(is_same_v<typename A::T, typename B::T>);
  using T = bool;
  static constexpr uint32_t fmask = A::fmask | B::fmask, imask = A::imask | B::imask, lmask = A::lmask | B::lmask;
  static constexpr bool uses_join = A::uses_join || B::uses_join;
  __device__ static __forceinline__ bool eval(const PairRegs& r, const Consts& c, int s) {
    if constexpr (is_same_v<typename A::T, double>) return cmp_f64(Op, A::eval(r, c, s), B::eval(r, c, s));
    else return cmp_int<typename A::T>(Op, A::eval(r, c, s), B::eval(r, c, s));
  }
#ifndef __CUDACC_RTC__
  static std::string sig() {
    static const char* names[] = {"lt", "lte", "gt", "gte", "eq", "neq"};
    return std::string(names[Op]) + "(" + A::sig() + "," + B::sig() + ")";
  }
#endif
};
template <class A, class B> using Lt = Compare<kLt, A, B>;
template <class A, class B> using Lte = Compare<kLte, A, B>;
template <class A, class B> using Gt = Compare<kGt, A, B>;
template <class A, class B> using Gte = Compare<kGte, A, B>;
template <class A, class B> using Eq = Compare<kEq, A, B>;
template <class A, class B> using Neq = Compare<kNeq, A, B>;

template <class X, class Lo, class Hi>
struct Between {
  using T = bool;
  static constexpr uint32_t fmask = X::fmask | Lo::fmask | Hi::fmask, imask = X::imask | Lo::imask | Hi::imask,
                            lmask = X::lmask | Lo::lmask | Hi::lmask;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ bool eval(const PairRegs& r, const Consts& c, int s) {
    const auto x = X::eval(r, c, s);
    if constexpr (is_same_v<typename X::T, double>) return gte_f64(x, Lo::eval(r, c, s)) && lte_f64(x, Hi::eval(r, c, s));
    else return x >= Lo::eval(r, c, s) && x <= Hi::eval(r, c, s);
  }
  VB2_SIG(static std::string sig() { return "between(" + X::sig() + "," + Lo::sig() + "," + Hi::sig() + ")"; })
};

template <class... As>
struct And {
  using T = bool;
  static constexpr uint32_t fmask = (As::fmask | ...), imask = (As::imask | ...), lmask = (As::lmask | ...);
  static constexpr bool uses_join = (As::uses_join || ...);
  // Null-free inputs: three-valued logic degenerates to &&. All conjuncts are evaluated (no
  // divergence); none of them can raise.
  __device__ static __forceinline__ bool eval(const PairRegs& r, const Consts& c, int s) { return (As::eval(r, c, s) & ...); }
#ifndef __CUDACC_RTC__
  static std::string sig() {
    std::string out = "and(";
    bool first = true;
    ((out += (first ? "" : ",") + As::sig(), first = false), ...);
    return out + ")";
  }
#endif
};

template <class C, class A, class B>
struct Switch {
  using T = typename A::T;
  static constexpr uint32_t fmask = C::fmask | A::fmask | B::fmask, imask = C::imask | A::imask | B::imask,
                            lmask = C::lmask | A::lmask | B::lmask;
  static constexpr bool uses_join = C::uses_join || A::uses_join || B::uses_join;
  __device__ static __forceinline__ T eval(const PairRegs& r, const Consts& c, int s) {
    return C::eval(r, c, s) ? A::eval(r, c, s) : B::eval(r, c, s);
  }
  VB2_SIG(static std::string sig() { return "switch(" + C::sig() + "," + A::sig() + "," + B::sig() + ")"; })
};

// ---- a pipeline = filter + projections (+ optional join probe on column JoinCol) ------------
template <class Filter, class Projs, int JoinCol = -1>
struct Pipeline;

template <class Filter, class... Ps, int JoinCol>
struct Pipeline<Filter, TypeList<Ps...>, JoinCol> {
  static constexpr int kNP = sizeof...(Ps);
  static constexpr bool kJoin = JoinCol >= 0;
  static constexpr int kJoinCol = JoinCol >= 0 ? JoinCol : 0;
  static constexpr uint32_t fmask = Filter::fmask | (Ps::fmask | ...);
  static constexpr uint32_t imask = Filter::imask | (Ps::imask | ...);
  static constexpr uint32_t lmask = Filter::lmask | (Ps::lmask | ...) | (kJoin ? (1u << (kJoin ? JoinCol : 0)) : 0u);
  static_assert(((is_same_v<typename Ps::T, double>) && ...), "fused projections are DOUBLE");
  using F = Filter;
  // The two halves of the late-materialisation path (selective filters): the filter alone, and
  // everything after it (join key + projections) for rows that already passed.
  struct FilterView {
    static constexpr uint32_t fmask = Filter::fmask, imask = Filter::imask, lmask = Filter::lmask;
    using F = Filter;
  };
  struct AfterFilter {
    static constexpr int kNP = sizeof...(Ps);
    static constexpr bool kJoin = JoinCol >= 0;
    static constexpr int kJoinCol = JoinCol >= 0 ? JoinCol : 0;
    static constexpr uint32_t fmask = (Ps::fmask | ...);
    static constexpr uint32_t imask = (Ps::imask | ...);
    static constexpr uint32_t lmask = (Ps::lmask | ...) | (kJoin ? (1u << (kJoin ? JoinCol : 0)) : 0u);
    using F = True;
    template <int I>
    __device__ static __forceinline__ void project(const PairRegs& r, const Consts& c, int s, double (&out)[kNP]) {
      Pipeline::template project<I>(r, c, s, out);
    }
  };
  template <int I>
  __device__ static __forceinline__ void project(const PairRegs& r, const Consts& c, int s, double (&out)[kNP]) {
    if constexpr (I < kNP) {
      using P = typename TypeAt<I, Ps...>::type;
      out[I] = P::eval(r, c, s);
      project<I + 1>(r, c, s, out);
    }
  }
#ifndef __CUDACC_RTC__
  static std::string sig() {
    std::string out = "F:" + Filter::sig() + ";P:";
    bool first = true;
    ((out += (first ? "" : "|") + Ps::sig(), first = false), ...);
    if (kJoin) out += ";J:l" + std::to_string(JoinCol);
    return out;
  }
#endif
};

struct KernelArgs {
  const void* cols[kMaxCols];
  Consts consts;
  int64_t rows;
  int32_t nkeys;
  int32_t ngroups;
  const void* key[VB2_FUSED_MAX_KEYS];
  int32_t key_is64[VB2_FUSED_MAX_KEYS];
  int32_t key_mult[VB2_FUSED_MAX_KEYS];
  int64_t key_min[VB2_FUSED_MAX_KEYS];
  const int32_t* key_lut[VB2_FUSED_MAX_KEYS];
  // Array-mode join table prepared for the fused probe: one byte per key slot,
  // 0 = no build row, 1 = match with build-side predicate false, 2 = match with predicate true.
  const uint8_t* join_slot_flags;
  int64_t join_min, join_range;
  // A value the fold of a tile's loaded registers is compared against before a stage is released
  // (see mbar_arrive_after); a runtime argument so that the comparison cannot be constant-folded.
  uint64_t release_guard;
};

constexpr uint64_t kReleaseGuard = 0x9e3779b97f4a7c15ull;
constexpr int kThreads = 256;

template <class P, bool kPair>
__device__ __forceinline__ void load_rows(const KernelArgs& a, int64_t idx, PairRegs& r) {
  // idx = pair index (kPair) or row index (!kPair; slot 0 only)
#pragma unroll
  for (int c = 0; c < kMaxCols; ++c) {
    if (P::fmask & (1u << c)) {
      if constexpr (kPair) {
        double2 v = ldg_stream_f64x2(reinterpret_cast<const double*>(a.cols[c]) + 2 * idx);
        r.f[c][0] = v.x; r.f[c][1] = v.y;
      } else {
        r.f[c][0] = __ldg(reinterpret_cast<const double*>(a.cols[c]) + idx);
      }
    }
    if (P::imask & (1u << c)) {
      if constexpr (kPair) {
        int2 v = __ldg(reinterpret_cast<const int2*>(a.cols[c]) + idx);
        r.i[c][0] = v.x; r.i[c][1] = v.y;
      } else {
        r.i[c][0] = __ldg(reinterpret_cast<const int32_t*>(a.cols[c]) + idx);
      }
    }
    if (P::lmask & (1u << c)) {
      if constexpr (kPair) {
        longlong2 v = ldg_stream_i64x2(reinterpret_cast<const int64_t*>(a.cols[c]) + 2 * idx);
        r.l[c][0] = v.x; r.l[c][1] = v.y;
      } else {
        r.l[c][0] = __ldg(reinterpret_cast<const int64_t*>(a.cols[c]) + idx);
      }
    }
  }
}

// Loads the group-key values of a pair/row (decoded into a group id later). KeyT = int32_t when
// every key column is 32-bit (dictionary indices, INTEGER, DATE), int64_t otherwise.
template <bool kPair, class KeyT>
__device__ __forceinline__ void load_keys(const KernelArgs& a, int64_t idx, KeyT (&kv)[VB2_FUSED_MAX_KEYS][2]) {
#pragma unroll
  for (int k = 0; k < VB2_FUSED_MAX_KEYS; ++k) {
    if (k < a.nkeys) {
      if (sizeof(KeyT) == 8 && a.key_is64[k]) {
        if constexpr (kPair) {
          longlong2 v = ldg_stream_i64x2(reinterpret_cast<const int64_t*>(a.key[k]) + 2 * idx);
          kv[k][0] = static_cast<KeyT>(v.x); kv[k][1] = static_cast<KeyT>(v.y);
        } else {
          kv[k][0] = static_cast<KeyT>(__ldg(reinterpret_cast<const int64_t*>(a.key[k]) + idx));
        }
      } else {
        if constexpr (kPair) {
          int2 v = __ldg(reinterpret_cast<const int2*>(a.key[k]) + idx);
          kv[k][0] = v.x; kv[k][1] = v.y;
        } else {
          kv[k][0] = __ldg(reinterpret_cast<const int32_t*>(a.key[k]) + idx);
        }
      }
    }
  }
}

template <class KeyT>
__device__ __forceinline__ int group_of(const KernelArgs& a, const KeyT (&kv)[VB2_FUSED_MAX_KEYS][2], int s) {
  int g = 0;
#pragma unroll
  for (int k = 0; k < VB2_FUSED_MAX_KEYS; ++k) {
    if (k < a.nkeys) {
      int64_t id = static_cast<int64_t>(kv[k][s]) - a.key_min[k];
      if (a.key_lut[k]) id = __ldg(a.key_lut[k] + id);
      g += static_cast<int>(id) * a.key_mult[k];
    }
  }
  return g;
}

template <class P, int kMaxG>
struct Accum {
  double sum[kMaxG][P::kNP];
  int32_t cnt[kMaxG];  // per-thread rows < 2^31; widened in the block reduction
  __device__ __forceinline__ void init() {
#pragma unroll
    for (int g = 0; g < kMaxG; ++g) {
      cnt[g] = 0;
#pragma unroll
      for (int p = 0; p < P::kNP; ++p) sum[g][p] = 0.0;
    }
  }
  // Branch-free: every accumulator stays in a register (no dynamic indexing, no divergence);
  // rows of other groups (and rows dropped by the filter, gid = -1) leave it untouched.
  __device__ __forceinline__ void add(int gid, const double (&v)[P::kNP]) {
#pragma unroll
    for (int g = 0; g < kMaxG; ++g) {
      const int hit = gid == g;
      cnt[g] += hit;
      // one predicated DADD per accumulator (ptxas shares the setp across the group's adds)
#pragma unroll
      for (int p = 0; p < P::kNP; ++p)
        asm("{ .reg .pred q; setp.ne.s32 q, %1, 0; @q add.rn.f64 %0, %0, %2; }" : "+d"(sum[g][p]) : "r"(hit), "d"(v[p]));
    }
  }
};

// Evaluates filter (+ join probe), group id and projections of one row slot.
template <class P, int kMaxG, class KeyT>
__device__ __forceinline__ void eval_slot(const KernelArgs& a, PairRegs& r, const KeyT (&kv)[VB2_FUSED_MAX_KEYS][2], int s,
                                          int& gid, double (&v)[P::kNP]) {
  bool keep = P::F::eval(r, a.consts, s);
  if constexpr (P::kJoin) {
    // Array-mode probe, branch-free: rows that failed the filter (or fall outside the key range)
    // read slot 0, which every lane shares, so the load costs no extra sectors.
    const int64_t slot = r.l[P::kJoinCol][s] - a.join_min;
    const bool in_range = keep && slot >= 0 && slot < a.join_range;
    const uint8_t hit = __ldg(a.join_slot_flags + (in_range ? slot : 0));
    keep = in_range && hit != 0;
    r.join_flag[s] = hit == 2;
  }
  P::template project<0>(r, a.consts, s, v);
  const int g = (kMaxG == 1) ? 0 : group_of<KeyT>(a, kv, s);
  gid = keep ? g : -1;
}

template <class P, int kMaxG, bool kPair, class KeyT>
__device__ __forceinline__ void process(const KernelArgs& a, const PairRegs& r0, const KeyT (&kv)[VB2_FUSED_MAX_KEYS][2],
                                        Accum<P, kMaxG>& acc) {
  PairRegs r = r0;
  int gid[2];
  double v[2][P::kNP];
#pragma unroll
  for (int s = 0; s < (kPair ? 2 : 1); ++s) eval_slot<P, kMaxG, KeyT>(a, r, kv, s, gid[s], v[s]);
#pragma unroll
  for (int s = 0; s < (kPair ? 2 : 1); ++s) acc.add(gid[s], v[s]);
}

// Block reduction: shuffles within warps, then across warps in fixed order; one partial per block.
constexpr int kMaxWarps = 9;
template <class P, int kMaxG>
__device__ __forceinline__ void block_reduce_store(const Accum<P, kMaxG>& acc, double* __restrict__ partials) {
  constexpr int kVals = kMaxG * (P::kNP + 1);
  __shared__ double red[kMaxWarps][kVals];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int nwarps = blockDim.x >> 5;
#pragma unroll
  for (int g = 0; g < kMaxG; ++g) {
#pragma unroll
    for (int p = 0; p < P::kNP; ++p) {
      double v = warp_sum(acc.sum[g][p]);
      if (lane == 0) red[warp][g * (P::kNP + 1) + p] = v;
    }
    int64_t c = warp_sum(static_cast<int64_t>(acc.cnt[g]));
    if (lane == 0) red[warp][g * (P::kNP + 1) + P::kNP] = __longlong_as_double(c);
  }
  __syncthreads();
  if (threadIdx.x < kVals) {
    const bool is_cnt = (threadIdx.x % (P::kNP + 1)) == P::kNP;
    double out;
    if (is_cnt) {
      int64_t c = 0;
      for (int w = 0; w < nwarps; ++w) c += __double_as_longlong(red[w][threadIdx.x]);
      out = __longlong_as_double(c);
    } else {
      double s = 0.0;
      for (int w = 0; w < nwarps; ++w) s = __dadd_rn(s, red[w][threadIdx.x]);
      out = s;
    }
    partials[static_cast<int64_t>(blockIdx.x) * kVals + threadIdx.x] = out;
  }
}

// Shared-memory accumulators for 5..kSmemMaxGroups groups: every consumer thread owns a private
// slot per (group, projection) at [(g * NP + p) * kConsumerThreads + tid] — dynamic group index,
// no atomics, no bank conflicts (a lane's bank pair depends on tid only). Cost per row is
// independent of the number of groups. Rows dropped by the filter go to a spare group.
template <class P>
struct SmemAccum {
  double* sums;   // [(G + 1) * NP][threads]
  int32_t* cnts;  // [(G + 1)][threads]
  int ngroups;    // G (spare group = index G)
  int threads;
  __device__ __forceinline__ void init(uint8_t* base, int g, int nthreads, int tid) {
    ngroups = g;
    threads = nthreads;
    sums = reinterpret_cast<double*>(base);
    cnts = reinterpret_cast<int32_t*>(base + static_cast<size_t>(g + 1) * P::kNP * nthreads * 8);
    for (int i = tid; i < (g + 1) * P::kNP * nthreads; i += nthreads) sums[i] = 0.0;
    for (int i = tid; i < (g + 1) * nthreads; i += nthreads) cnts[i] = 0;
  }
  static __host__ __device__ size_t bytes(int g, int nthreads) { return static_cast<size_t>(g + 1) * nthreads * (P::kNP * 8 + 4); }
  __device__ __forceinline__ void add(int gid, const double (&v)[P::kNP], int tid) {
    const int g = gid < 0 ? ngroups : gid;
    double* s = sums + static_cast<size_t>(g) * P::kNP * threads + tid;
#pragma unroll
    for (int p = 0; p < P::kNP; ++p) s[p * threads] = __dadd_rn(s[p * threads], v[p]);
    cnts[g * threads + tid] += 1;
  }
  // one partial per block: kvals = G * (NP + 1), fixed reduction order over the threads
  __device__ __forceinline__ void reduce_store(double* __restrict__ partials, int kvals) {
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int t = warp; t < kvals; t += nwarps) {
      const int g = t / (P::kNP + 1), p = t % (P::kNP + 1);
      double out;
      if (p == P::kNP) {
        int64_t c = 0;
        for (int i = lane; i < threads; i += 32) c += cnts[g * threads + i];
        out = __longlong_as_double(warp_sum(c));
      } else {
        double x = 0.0;
        for (int i = lane; i < threads; i += 32) x = __dadd_rn(x, sums[(static_cast<size_t>(g) * P::kNP + p) * threads + i]);
        out = warp_sum(x);
      }
      if (lane == 0) partials[static_cast<int64_t>(blockIdx.x) * kvals + t] = out;
    }
  }
};

template <class P, int kMaxG, int kUnroll, bool kPair, class KeyT>
__global__ void __launch_bounds__(kThreads, kMaxG <= 4 ? 2 : 1) fused_scan_agg_kernel(const __grid_constant__ KernelArgs a, double* __restrict__ partials) {
  Accum<P, kMaxG> acc;
  acc.init();
  const int64_t units = kPair ? (a.rows >> 1) : a.rows;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  int64_t u0 = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x;
  // main loop: kUnroll independent units in flight per thread
  for (; u0 + (kUnroll - 1) * stride < units; u0 += kUnroll * stride) {
    PairRegs r[kUnroll];
    KeyT kv[kUnroll][VB2_FUSED_MAX_KEYS][2];
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) {
      load_rows<P, kPair>(a, u0 + j * stride, r[j]);
      if (kMaxG > 1) load_keys<kPair, KeyT>(a, u0 + j * stride, kv[j]);
    }
#pragma unroll
    for (int j = 0; j < kUnroll; ++j) process<P, kMaxG, kPair, KeyT>(a, r[j], kv[j], acc);
  }
  for (; u0 < units; u0 += stride) {
    PairRegs r;
    KeyT kv[VB2_FUSED_MAX_KEYS][2];
    load_rows<P, kPair>(a, u0, r);
    if (kMaxG > 1) load_keys<kPair, KeyT>(a, u0, kv);
    process<P, kMaxG, kPair, KeyT>(a, r, kv, acc);
  }
  if (kPair && (a.rows & 1) && blockIdx.x == 0 && threadIdx.x == 0) {  // odd tail row
    PairRegs r;
    KeyT kv[VB2_FUSED_MAX_KEYS][2];
    load_rows<P, false>(a, a.rows - 1, r);
    if (kMaxG > 1) load_keys<false, KeyT>(a, a.rows - 1, kv);
    process<P, kMaxG, false, KeyT>(a, r, kv, acc);
  }
  block_reduce_store<P, kMaxG>(acc, partials);
}

// ---------------------------------------------------------------------------------------------
// TMA-staged variant (main path). One producer warp streams kTileRows-row slices of every
// referenced column into a ring of shared-memory stages with 1-D bulk async copies
// (cp.async.bulk ... mbarrier::complete_tx), eight consumer warps evaluate the pipeline out of
// shared memory. The bytes in flight per SM are stages x tile bytes (~90-180 KB) regardless of
// how ptxas schedules the consumers' code, which is what an HBM-bound scan needs.
// ---------------------------------------------------------------------------------------------
constexpr int kTileRows = 1024;
constexpr int kConsumerThreads = 256;
constexpr int kTmaThreads = kConsumerThreads + 32;
constexpr int kMaxStages = 4;
constexpr int kRowsPerThread = kTileRows / kConsumerThreads;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Releases a stage to the producer once the values loaded from it are in registers. Measured on
// B200: an ld.shared that was merely *issued* before mbarrier.arrive can still observe the
// producer's next bulk copy into the stage (a pipeline whose last load result was consumed after
// the arrive read stale tiles a few times per thousand). So the arrive is made data-dependent on
// every loaded value: `dep` folds the loaded registers and the arrive is predicated on comparing it
// with a runtime guard value — the compare cannot issue before all of the thread's ld.shared
// results have arrived (register scoreboard), and it cannot be folded away. Both outcomes arrive.
__device__ __forceinline__ void mbar_arrive_after(uint64_t* bar, uint64_t dep, uint64_t guard) {
  asm volatile(
      "{ .reg .pred q;\n"
      "  setp.ne.b64 q, %1, %2;\n"
      "  @q mbarrier.arrive.shared::cta.b64 _, [%0];\n"
      "  @!q mbarrier.arrive.shared::cta.b64 _, [%0], 1; }"
      ::"r"(smem_u32(bar)), "l"(dep), "l"(guard)
      : "memory");
}
__device__ __forceinline__ uint64_t warp_xor(uint64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v ^= __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {}
}
// 1-D bulk async copy global -> shared; completion is signalled on `bar` as transaction bytes.
__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// Byte offsets of the referenced columns inside one stage: f64 columns, i64 columns, 8-byte keys,
// i32 columns, 4-byte keys. Everything but the key count is a compile-time constant.
template <class P, int kKeyBytes, int kTileRows = ::vb2::fx::kTileRows>
struct TileLayout {
  __host__ __device__ static constexpr int popc(uint32_t m) { int n = 0; for (; m; m &= m - 1) ++n; return n; }
  __host__ __device__ static constexpr uint32_t low(int c) { return (1u << c) - 1u; }
  static constexpr int kBase8 = 8 * kTileRows * (popc(P::fmask) + popc(P::lmask));
  __host__ __device__ static constexpr int f_off(int c) { return 8 * kTileRows * popc(P::fmask & low(c)); }
  __host__ __device__ static constexpr int l_off(int c) { return 8 * kTileRows * (popc(P::fmask) + popc(P::lmask & low(c))); }
  __host__ __device__ static constexpr int i_base(int nkeys) { return kBase8 + (kKeyBytes == 8 ? nkeys * 8 * kTileRows : 0); }
  __host__ __device__ static constexpr int i_off(int c, int nkeys) { return i_base(nkeys) + 4 * kTileRows * popc(P::imask & low(c)); }
  __host__ __device__ static constexpr int key_off(int k, int nkeys) {
    return kKeyBytes == 8 ? kBase8 + 8 * kTileRows * k : i_base(nkeys) + 4 * kTileRows * (popc(P::imask) + k);
  }
  __host__ __device__ static constexpr int stage_bytes(int nkeys) {
    const int end = i_base(nkeys) + 4 * kTileRows * popc(P::imask) + (kKeyBytes == 4 ? nkeys * 4 * kTileRows : 0);
    return (end + 127) / 128 * 128;
  }
};

// kMaxG > 0: register accumulators for up to kMaxG groups; kMaxG == 0: shared-memory accumulators.
template <class P, int kMaxG, class KeyT>
__global__ void __launch_bounds__(kTmaThreads, kMaxG == 0 ? 1 : 2)
fused_scan_agg_tma_kernel(const __grid_constant__ KernelArgs a, int stages, double* __restrict__ partials) {
  extern __shared__ __align__(128) uint8_t tile_smem[];
  constexpr bool kSmemAcc = kMaxG == 0;
  constexpr int kRegG = kSmemAcc ? 1 : kMaxG;
  constexpr bool kHasKeys = kSmemAcc || kMaxG > 1;
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  using Lay = TileLayout<P, sizeof(KeyT)>;
  const int nk = kHasKeys ? a.nkeys : 0;
  const int stage_bytes = Lay::stage_bytes(nk);
  const int64_t ntiles = a.rows / kTileRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);                        // producer's expect_tx arrival
      mbar_init(&empty_bar[s], kConsumerThreads);  // one arrival per consumer thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  Accum<P, kRegG> acc;
  acc.init();
  SmemAccum<P> sacc;
  if constexpr (kSmemAcc) {
    // accumulators live behind the tile stages; the producer warp helps zeroing them
    sacc.init(tile_smem + static_cast<size_t>(stages) * stage_bytes, a.ngroups, kConsumerThreads, threadIdx.x);
    sacc.threads = kConsumerThreads;
  }
  __syncthreads();
  if (warp == kConsumerThreads / kWarp) {
    // ---- producer warp: one elected lane issues the copies ----
    if (lane == 0) {
      int it = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % stages;
        const uint32_t round = static_cast<uint32_t>(it / stages);
        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
        uint8_t* base = tile_smem + static_cast<size_t>(s) * stage_bytes;
        const int64_t row0 = t * kTileRows;
        uint32_t bytes = 0;
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) {
          if (P::fmask & (1u << c)) bytes += kTileRows * 8;
          if (P::lmask & (1u << c)) bytes += kTileRows * 8;
          if (P::imask & (1u << c)) bytes += kTileRows * 4;
        }
        if (kHasKeys) bytes += a.nkeys * kTileRows * sizeof(KeyT);
        mbar_expect_tx(&full_bar[s], bytes);
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) {
          if (P::fmask & (1u << c)) bulk_load(base + Lay::f_off(c), reinterpret_cast<const double*>(a.cols[c]) + row0, kTileRows * 8, &full_bar[s]);
          if (P::lmask & (1u << c)) bulk_load(base + Lay::l_off(c), reinterpret_cast<const int64_t*>(a.cols[c]) + row0, kTileRows * 8, &full_bar[s]);
          if (P::imask & (1u << c)) bulk_load(base + Lay::i_off(c, nk), reinterpret_cast<const int32_t*>(a.cols[c]) + row0, kTileRows * 4, &full_bar[s]);
        }
        if (kHasKeys) {
          for (int k = 0; k < a.nkeys; ++k)
            bulk_load(base + Lay::key_off(k, nk), reinterpret_cast<const KeyT*>(a.key[k]) + row0, kTileRows * sizeof(KeyT), &full_bar[s]);
        }
      }
    }
  } else {
    // ---- consumer warps ----
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % stages;
      const uint32_t round = static_cast<uint32_t>(it / stages);
      mbar_wait(&full_bar[s], round & 1);
      const uint8_t* base = tile_smem + static_cast<size_t>(s) * stage_bytes;
      PairRegs r[kRowsPerThread];
      KeyT kv[kRowsPerThread][VB2_FUSED_MAX_KEYS][2];
      uint64_t dep = 0;
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j) {
        const int row = j * kConsumerThreads + threadIdx.x;
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) {
          if (P::fmask & (1u << c)) {
            r[j].f[c][0] = reinterpret_cast<const double*>(base + Lay::f_off(c))[row];
            dep ^= static_cast<uint64_t>(__double_as_longlong(r[j].f[c][0]));
          }
          if (P::lmask & (1u << c)) {
            r[j].l[c][0] = reinterpret_cast<const int64_t*>(base + Lay::l_off(c))[row];
            dep ^= static_cast<uint64_t>(r[j].l[c][0]);
          }
          if (P::imask & (1u << c)) {
            r[j].i[c][0] = reinterpret_cast<const int32_t*>(base + Lay::i_off(c, nk))[row];
            dep ^= static_cast<uint64_t>(static_cast<uint32_t>(r[j].i[c][0]));
          }
        }
        if (kHasKeys) {
#pragma unroll
          for (int k = 0; k < VB2_FUSED_MAX_KEYS; ++k)
            if (k < a.nkeys) {
              kv[j][k][0] = reinterpret_cast<const KeyT*>(base + Lay::key_off(k, nk))[row];
              dep ^= static_cast<uint64_t>(kv[j][k][0]);
            }
        }
      }
      // every value of this stage is in a register: hand it back to the producer before computing.
      // Every consumer thread arrives for itself (barrier count = consumer threads), predicated on
      // the fold of ITS OWN loads: no lane's release can overtake that lane's reads of the stage.
      mbar_arrive_after(&empty_bar[s], dep, a.release_guard);
      int gid[kRowsPerThread];
      double v[kRowsPerThread][P::kNP];
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j) eval_slot<P, kHasKeys ? 2 : 1, KeyT>(a, r[j], kv[j], 0, gid[j], v[j]);
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j) {
        if constexpr (kSmemAcc) sacc.add(gid[j], v[j], threadIdx.x);
        else acc.add(gid[j], v[j]);
      }
    }
    // tail rows (rows % kTileRows) by direct loads, spread over the blocks' consumer threads
    const int64_t tail0 = ntiles * kTileRows;
    for (int64_t row = tail0 + static_cast<int64_t>(blockIdx.x) * kConsumerThreads + threadIdx.x; row < a.rows;
         row += static_cast<int64_t>(gridDim.x) * kConsumerThreads) {
      PairRegs r;
      KeyT kv[VB2_FUSED_MAX_KEYS][2];
      load_rows<P, false>(a, row, r);
      if (kHasKeys) load_keys<false, KeyT>(a, row, kv);
      int gid;
      double v[P::kNP];
      eval_slot<P, kHasKeys ? 2 : 1, KeyT>(a, r, kv, 0, gid, v);
      if constexpr (kSmemAcc) sacc.add(gid, v, threadIdx.x);
      else acc.add(gid, v);
    }
  }
  if constexpr (kSmemAcc) sacc.reduce_store(partials, a.ngroups * (P::kNP + 1));
  else block_reduce_store<P, kRegG>(acc, partials);
}

// ---------------------------------------------------------------------------------------------
// Fused scan -> filter -> project -> compact: same TMA-staged producer, the consumers write the
// projections of the surviving rows densely. Used in front of the hash-partitioned exchange
// (each GPU filters its lineitem shard and ships only the ~1 % of rows that join). One atomic
// per tile reserves the output range; order inside a tile follows the input, tiles land in
// arrival order (the exchange that follows is a multiset operation).
// ---------------------------------------------------------------------------------------------
template <class Filter, class Projs>
struct CompactPipeline;
template <class Filter, class... Ps>
struct CompactPipeline<Filter, TypeList<Ps...>> {
  static constexpr int kNP = sizeof...(Ps);
  static constexpr bool kJoin = false;
  static constexpr int kJoinCol = 0;
  static constexpr uint32_t fmask = Filter::fmask | (Ps::fmask | ...);
  static constexpr uint32_t imask = Filter::imask | (Ps::imask | ...);
  static constexpr uint32_t lmask = Filter::lmask | (Ps::lmask | ...);
  using F = Filter;
  template <int I>
  __device__ static __forceinline__ void store(const PairRegs& r, const Consts& c, void* const* outs, int64_t pos) {
    if constexpr (I < kNP) {
      using P = typename TypeAt<I, Ps...>::type;
      reinterpret_cast<typename P::T*>(outs[I])[pos] = P::eval(r, c, 0);
      store<I + 1>(r, c, outs, pos);
    }
  }
#ifndef __CUDACC_RTC__
  static std::string sig() {
    std::string out = "F:" + Filter::sig() + ";C:";
    bool first = true;
    ((out += (first ? "" : "|") + Ps::sig(), first = false), ...);
    return out;
  }
#endif
#ifndef __CUDACC_RTC__
  static void widths(int* w) {
    int i = 0;
    ((w[i++] = static_cast<int>(sizeof(typename Ps::T))), ...);
  }
#endif
};

struct CompactArgs {
  void* outs[VB2_FUSED_MAX_COLS];
  int64_t capacity;
  unsigned long long* count;  // device counter of rows written
  int32_t* error_flag;        // set to 100 when capacity is exceeded
};

template <class P>
__global__ void __launch_bounds__(kTmaThreads, 2)
fused_scan_compact_tma_kernel(const __grid_constant__ KernelArgs a, const __grid_constant__ CompactArgs o, int stages) {
  extern __shared__ __align__(128) uint8_t tile_smem[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  __shared__ int warp_counts[kConsumerThreads / kWarp];
  __shared__ long long tile_base;
  using Lay = TileLayout<P, 4>;
  const int stage_bytes = Lay::stage_bytes(0);
  const int64_t ntiles = a.rows / kTileRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerThreads);  // one arrival per consumer thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (warp == kConsumerThreads / kWarp) {
    if (lane == 0) {
      int it = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % stages;
        const uint32_t round = static_cast<uint32_t>(it / stages);
        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
        uint8_t* base = tile_smem + static_cast<size_t>(s) * stage_bytes;
        const int64_t row0 = t * kTileRows;
        uint32_t bytes = 0;
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) {
          if (P::fmask & (1u << c)) bytes += kTileRows * 8;
          if (P::lmask & (1u << c)) bytes += kTileRows * 8;
          if (P::imask & (1u << c)) bytes += kTileRows * 4;
        }
        mbar_expect_tx(&full_bar[s], bytes);
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) {
          if (P::fmask & (1u << c)) bulk_load(base + Lay::f_off(c), reinterpret_cast<const double*>(a.cols[c]) + row0, kTileRows * 8, &full_bar[s]);
          if (P::lmask & (1u << c)) bulk_load(base + Lay::l_off(c), reinterpret_cast<const int64_t*>(a.cols[c]) + row0, kTileRows * 8, &full_bar[s]);
          if (P::imask & (1u << c)) bulk_load(base + Lay::i_off(c, 0), reinterpret_cast<const int32_t*>(a.cols[c]) + row0, kTileRows * 4, &full_bar[s]);
        }
      }
    }
  } else {
    int it = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
      const int s = it % stages;
      const uint32_t round = static_cast<uint32_t>(it / stages);
      mbar_wait(&full_bar[s], round & 1);
      const uint8_t* base = tile_smem + static_cast<size_t>(s) * stage_bytes;
      PairRegs r[kRowsPerThread];
      bool keep[kRowsPerThread];
      int before = 0;  // kept rows of this warp that precede this lane's rows, per slot order
      int mine[kRowsPerThread];
      int warp_total = 0;
      uint64_t dep = 0;
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j) {
        // row order inside the tile: warp-major so that a warp owns 128 consecutive rows
        const int row = warp * (kWarp * kRowsPerThread) + j * kWarp + lane;
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) {
          if (P::fmask & (1u << c)) {
            r[j].f[c][0] = reinterpret_cast<const double*>(base + Lay::f_off(c))[row];
            dep ^= static_cast<uint64_t>(__double_as_longlong(r[j].f[c][0]));
          }
          if (P::lmask & (1u << c)) {
            r[j].l[c][0] = reinterpret_cast<const int64_t*>(base + Lay::l_off(c))[row];
            dep ^= static_cast<uint64_t>(r[j].l[c][0]);
          }
          if (P::imask & (1u << c)) {
            r[j].i[c][0] = reinterpret_cast<const int32_t*>(base + Lay::i_off(c, 0))[row];
            dep ^= static_cast<uint64_t>(static_cast<uint32_t>(r[j].i[c][0]));
          }
        }
        keep[j] = P::F::eval(r[j], a.consts, 0);
        const unsigned m = __ballot_sync(0xffffffffu, keep[j]);
        mine[j] = warp_total + __popc(m & ((1u << lane) - 1));
        warp_total += __popc(m);
      }
      (void)before;
      mbar_arrive_after(&empty_bar[s], dep, a.release_guard);
      if (lane == 0) warp_counts[warp] = warp_total;
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
      if (threadIdx.x == 0) {
        int total = 0;
        for (int w = 0; w < kConsumerThreads / kWarp; ++w) total += warp_counts[w];
        tile_base = total ? static_cast<long long>(atomicAdd(o.count, static_cast<unsigned long long>(total))) : 0;
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");
      int64_t pos0 = tile_base;
      for (int w = 0; w < warp; ++w) pos0 += warp_counts[w];
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j) {
        if (keep[j]) {
          const int64_t pos = pos0 + mine[j];
          if (pos < o.capacity) P::template store<0>(r[j], a.consts, o.outs, pos);
          else atomicCAS(o.error_flag, 0, 100);
        }
      }
      asm volatile("bar.sync 1, %0;" ::"n"(kConsumerThreads) : "memory");  // warp_counts reused next tile
    }
    // tail rows by direct loads (one atomic per kept row; at most kTileRows - 1 rows)
    const int64_t tail0 = ntiles * kTileRows;
    for (int64_t row = tail0 + static_cast<int64_t>(blockIdx.x) * kConsumerThreads + threadIdx.x; row < a.rows;
         row += static_cast<int64_t>(gridDim.x) * kConsumerThreads) {
      PairRegs r;
      load_rows<P, false>(a, row, r);
      if (P::F::eval(r, a.consts, 0)) {
        const int64_t pos = static_cast<int64_t>(atomicAdd(o.count, 1ull));
        if (pos < o.capacity) P::template store<0>(r, a.consts, o.outs, pos);
        else atomicCAS(o.error_flag, 0, 100);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Late materialisation for selective filters (the reference's FilterProject does the same thing
// row-set-wise: the filter runs first and the projections only see the surviving rows,
// velox/exec/FilterProject.cpp:200-259; its readers go further and load the other columns lazily).
//   fused_filter_bits_tma_kernel   streams ONLY the filter's columns through the TMA ring and writes
//                                  the selection bitmap (1 bit per row, 32-bit words); with
//                                  tile_stride > 1 it visits every tile_stride-th tile and only
//                                  counts (selectivity estimate for the planner).
//   fused_gather_agg_kernel        join probe + projections + aggregation over the selected row
//                                  numbers: columns are touched at the sectors of surviving rows only.
// Q14 keeps 1.2 % of lineitem: 4 B/row of filter traffic + ~5 % of the other columns' sectors
// instead of 28 B/row.
// ---------------------------------------------------------------------------------------------
__host__ __device__ constexpr int filter_tile_rows_for(uint32_t fmask, uint32_t imask, uint32_t lmask) {
  int bytes = 0;
  for (int c = 0; c < 32; ++c) bytes += ((fmask >> c) & 1u) * 8 + ((lmask >> c) & 1u) * 8 + ((imask >> c) & 1u) * 4;
  return bytes <= 4 ? 4096 : (bytes <= 8 ? 2048 : 1024);
}
// kFilterTileRows: rows per tile — a filter over one 4-byte column moves only 4 KB per 1024 rows, too
// little to amortise a stage hand-off, so narrow filters take taller tiles (>= 16 KB per stage).
template <class FV, int kFilterTileRows>
__global__ void __launch_bounds__(kTmaThreads, 2)
fused_filter_bits_tma_kernel(const __grid_constant__ KernelArgs a, int stages, int tile_stride, uint32_t* __restrict__ bits,
                             unsigned long long* __restrict__ counters) {
  extern __shared__ __align__(128) uint8_t tile_smem[];
  __shared__ uint64_t full_bar[kMaxStages], empty_bar[kMaxStages];
  constexpr int kTileRows = kFilterTileRows;  // shadows the pipeline-wide constant inside this kernel
  constexpr int kRowsPerThread = kFilterTileRows / kConsumerThreads;
  using Lay = TileLayout<FV, 4, kFilterTileRows>;
  const int stage_bytes = Lay::stage_bytes(0);
  const int64_t ntiles = a.rows / kTileRows;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], kConsumerThreads);  // one arrival per consumer thread
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t first = static_cast<int64_t>(blockIdx.x) * tile_stride, step = static_cast<int64_t>(gridDim.x) * tile_stride;
  if (warp == kConsumerThreads / kWarp) {
    if (lane == 0) {
      int it = 0;
      for (int64_t t = first; t < ntiles; t += step, ++it) {
        const int s = it % stages;
        const uint32_t round = static_cast<uint32_t>(it / stages);
        if (round > 0) mbar_wait(&empty_bar[s], (round - 1) & 1);
        uint8_t* base = tile_smem + static_cast<size_t>(s) * stage_bytes;
        const int64_t row0 = t * kTileRows;
        uint32_t bytes = 0;
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) {
          if (FV::fmask & (1u << c)) bytes += kTileRows * 8;
          if (FV::lmask & (1u << c)) bytes += kTileRows * 8;
          if (FV::imask & (1u << c)) bytes += kTileRows * 4;
        }
        mbar_expect_tx(&full_bar[s], bytes);
#pragma unroll
        for (int c = 0; c < kMaxCols; ++c) {
          if (FV::fmask & (1u << c)) bulk_load(base + Lay::f_off(c), reinterpret_cast<const double*>(a.cols[c]) + row0, kTileRows * 8, &full_bar[s]);
          if (FV::lmask & (1u << c)) bulk_load(base + Lay::l_off(c), reinterpret_cast<const int64_t*>(a.cols[c]) + row0, kTileRows * 8, &full_bar[s]);
          if (FV::imask & (1u << c)) bulk_load(base + Lay::i_off(c, 0), reinterpret_cast<const int32_t*>(a.cols[c]) + row0, kTileRows * 4, &full_bar[s]);
        }
      }
    }
    return;
  }
  unsigned long long kept = 0, seen = 0;
  int it = 0;
  for (int64_t t = first; t < ntiles; t += step, ++it) {
    const int s = it % stages;
    const uint32_t round = static_cast<uint32_t>(it / stages);
    mbar_wait(&full_bar[s], round & 1);
    const uint8_t* base = tile_smem + static_cast<size_t>(s) * stage_bytes;
    PairRegs r[kRowsPerThread];
    uint64_t dep = 0;
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      const int row = j * kConsumerThreads + threadIdx.x;  // a warp owns 32 consecutive rows: one bitmap word
#pragma unroll
      for (int c = 0; c < kMaxCols; ++c) {
        if (FV::fmask & (1u << c)) {
          r[j].f[c][0] = reinterpret_cast<const double*>(base + Lay::f_off(c))[row];
          dep ^= static_cast<uint64_t>(__double_as_longlong(r[j].f[c][0]));
        }
        if (FV::lmask & (1u << c)) {
          r[j].l[c][0] = reinterpret_cast<const int64_t*>(base + Lay::l_off(c))[row];
          dep ^= static_cast<uint64_t>(r[j].l[c][0]);
        }
        if (FV::imask & (1u << c)) {
          r[j].i[c][0] = reinterpret_cast<const int32_t*>(base + Lay::i_off(c, 0))[row];
          dep ^= static_cast<uint64_t>(static_cast<uint32_t>(r[j].i[c][0]));
        }
      }
    }
    mbar_arrive_after(&empty_bar[s], dep, a.release_guard);
#pragma unroll
    for (int j = 0; j < kRowsPerThread; ++j) {
      const bool keep = FV::F::eval(r[j], a.consts, 0);
      const unsigned word = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) {
        if (bits) bits[(t * kTileRows + j * kConsumerThreads + warp * kWarp) >> 5] = word;
        kept += __popc(word);
        seen += kWarp;
      }
    }
  }
  // tail rows (rows % kTileRows): whole bitmap words by direct loads, block 0 only (full scans only)
  if (tile_stride == 1 && blockIdx.x == 0) {
    const int64_t tail0 = ntiles * kTileRows;
    for (int64_t w0 = tail0 + static_cast<int64_t>(warp) * kWarp; w0 < a.rows; w0 += kConsumerThreads) {
      const int64_t row = w0 + lane;
      bool keep = false;
      if (row < a.rows) {
        PairRegs r;
        load_rows<FV, false>(a, row, r);
        keep = FV::F::eval(r, a.consts, 0);
      }
      const unsigned word = __ballot_sync(0xffffffffu, keep);
      if (lane == 0) {
        if (bits) bits[w0 >> 5] = word;
        kept += __popc(word);
        seen += (a.rows - w0 < kWarp) ? (a.rows - w0) : kWarp;
      }
    }
  }
  if (lane == 0 && counters && seen) {
    atomicAdd(counters, kept);
    atomicAdd(counters + 1, seen);
  }
}

// Rows sel[0 .. nsel) (ascending row numbers that passed the filter): probe, project, aggregate.
// Register accumulators (<= kMaxG groups); four independent rows in flight per thread.
template <class P, int kMaxG, class KeyT>
__global__ void __launch_bounds__(kThreads, 2)
fused_gather_agg_kernel(const __grid_constant__ KernelArgs a, const int32_t* __restrict__ sel, const int64_t* __restrict__ nsel_dev,
                        double* __restrict__ partials) {
  using AF = typename P::AfterFilter;
  Accum<AF, kMaxG> acc;
  acc.init();
  constexpr int kU = 4;
  const int64_t nsel = *nsel_dev;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * kThreads;
  for (int64_t i0 = static_cast<int64_t>(blockIdx.x) * kThreads + threadIdx.x; i0 < nsel; i0 += kU * stride) {
    PairRegs r[kU];
    KeyT kv[kU][VB2_FUSED_MAX_KEYS][2];
    bool ok[kU];
    int64_t row[kU];
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      const int64_t i = i0 + j * stride;
      ok[j] = i < nsel;
      row[j] = ok[j] ? sel[i] : sel[i0];
    }
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      load_rows<AF, false>(a, row[j], r[j]);
      if (kMaxG > 1) load_keys<false, KeyT>(a, row[j], kv[j]);
    }
#pragma unroll
    for (int j = 0; j < kU; ++j) {
      int gid;
      double v[AF::kNP];
      eval_slot<AF, kMaxG, KeyT>(a, r[j], kv[j], 0, gid, v);
      acc.add(ok[j] ? gid : -1, v);
    }
  }
  block_reduce_store<AF, kMaxG>(acc, partials);
}

#ifndef __CUDACC_RTC__
// Folds per-block partials in block order into the persistent accumulators.
__global__ void fused_finalize_kernel(const double* __restrict__ partials, int nblocks, int kvals, int np, int maxg,
                                      int ngroups, double* __restrict__ sums, int64_t* __restrict__ counts);

// Which kernels of a pipeline exist is the same for ahead-of-time template instantiations and for
// NVRTC-compiled ones (fused_jit.cu); the launch logic below (fused_scan.cu) only sees this interface.
enum class KernelKind : int { kTma = 0, kDirect = 1, kFilterBits = 2, kGather = 3 };
struct PipelineDesc {
  int nproj = 0;
  bool join = false;
  bool has_filter = false;                     // a filter other than `true`: the late-materialisation kernels exist
  uint32_t fmask = 0, imask = 0, lmask = 0;      // columns of the whole pipeline by type
  uint32_t ffmask = 0, fimask = 0, flmask = 0;   // columns the filter alone reads
};
// Kernel entry point for (kind, accumulator variant, key width), or nullptr. max_groups: 1 / 4 / 8
// register accumulators, 0 = shared-memory accumulators (kTma only); ignored for kFilterBits.
using KernelGetter = const void* (*)(void* self, KernelKind kind, int max_groups, bool key64);
using CompactFn = int (*)(const KernelArgs&, const CompactArgs&, cudaStream_t);

struct Entry {
  std::string signature;
  int nproj = 0;
  bool join = false;
  PipelineDesc desc;
  KernelGetter kernels = nullptr;  // aggregate pipelines (and filter-only pipelines: nproj == 0)
  void* self = nullptr;
  CompactFn compact = nullptr;     // compaction pipelines (";C:" signatures)
  int widths[VB2_FUSED_MAX_COLS] = {0};
};

int register_pipeline(const Entry& e);
#endif  // __CUDACC_RTC__

}  // namespace fx
}  // namespace vb2
