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
#include <string>
#include <tuple>
#include <type_traits>

#include "common.cuh"

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

// ---- leaves ---------------------------------------------------------------------------------
template <int C>
struct ColF {
  using T = double;
  static constexpr uint32_t fmask = 1u << C, imask = 0, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ double eval(const PairRegs& r, const Consts&, int s) { return r.f[C][s]; }
  static std::string sig() { return "f" + std::to_string(C); }
};
template <int C>
struct ColI {
  using T = int32_t;
  static constexpr uint32_t fmask = 0, imask = 1u << C, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ int32_t eval(const PairRegs& r, const Consts&, int s) { return r.i[C][s]; }
  static std::string sig() { return "i" + std::to_string(C); }
};
template <int C>
struct ColL {
  using T = int64_t;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 1u << C;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ int64_t eval(const PairRegs& r, const Consts&, int s) { return r.l[C][s]; }
  static std::string sig() { return "l" + std::to_string(C); }
};
template <int K>
struct PF {
  using T = double;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ double eval(const PairRegs&, const Consts& c, int) { return c.pf[K]; }
  static std::string sig() { return "pf" + std::to_string(K); }
};
template <int K>
struct PI {
  using T = int32_t;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ int32_t eval(const PairRegs&, const Consts& c, int) { return c.pi[K]; }
  static std::string sig() { return "pi" + std::to_string(K); }
};
template <int K>
struct PL {
  using T = int64_t;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ int64_t eval(const PairRegs&, const Consts& c, int) { return c.pl[K]; }
  static std::string sig() { return "pl" + std::to_string(K); }
};
struct True {
  using T = bool;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 0;
  static constexpr bool uses_join = false;
  __device__ static __forceinline__ bool eval(const PairRegs&, const Consts&, int) { return true; }
  static std::string sig() { return "true"; }
};
// Build-side predicate of the matched build row (e.g. p_type LIKE 'PROMO%'), evaluated once per
// dictionary entry on the build side and looked up here.
struct JoinFlag {
  using T = bool;
  static constexpr uint32_t fmask = 0, imask = 0, lmask = 0;
  static constexpr bool uses_join = true;
  __device__ static __forceinline__ bool eval(const PairRegs& r, const Consts&, int s) { return r.join_flag[s]; }
  static std::string sig() { return "joinflag"; }
};

// ---- operators ------------------------------------------------------------------------------
#define VB2_FX_BINARY(Name, text, expr)                                                          \
  template <class A, class B>                                                                    \
  struct Name {                                                                                  \
    static_assert(std::is_same_v<typename A::T, double> && std::is_same_v<typename B::T, double>); \
    using T = double;                                                                            \
    static constexpr uint32_t fmask = A::fmask | B::fmask, imask = A::imask | B::imask,          \
                              lmask = A::lmask | B::lmask;                                       \
    static constexpr bool uses_join = A::uses_join || B::uses_join;                              \
    __device__ static __forceinline__ double eval(const PairRegs& r, const Consts& c, int s) {   \
      const double a = A::eval(r, c, s), b = B::eval(r, c, s);                                   \
      return expr;                                                                               \
    }                                                                                            \
    static std::string sig() { return std::string(text "(") + A::sig() + "," + B::sig() + ")"; } \
  };
VB2_FX_BINARY(Plus, "plus", __dadd_rn(a, b))
VB2_FX_BINARY(Minus, "minus", __dsub_rn(a, b))
VB2_FX_BINARY(Multiply, "multiply", __dmul_rn(a, b))
VB2_FX_BINARY(Divide, "divide", __ddiv_rn(a, b))
#undef VB2_FX_BINARY

template <int Op, class A, class B>
struct Compare {
  static_assert(std::is_same_v<typename A::T, typename B::T>);
  using T = bool;
  static constexpr uint32_t fmask = A::fmask | B::fmask, imask = A::imask | B::imask, lmask = A::lmask | B::lmask;
  static constexpr bool uses_join = A::uses_join || B::uses_join;
  __device__ static __forceinline__ bool eval(const PairRegs& r, const Consts& c, int s) {
    if constexpr (std::is_same_v<typename A::T, double>) return cmp_f64(Op, A::eval(r, c, s), B::eval(r, c, s));
    else return cmp_int<typename A::T>(Op, A::eval(r, c, s), B::eval(r, c, s));
  }
  static std::string sig() {
    static const char* names[] = {"lt", "lte", "gt", "gte", "eq", "neq"};
    return std::string(names[Op]) + "(" + A::sig() + "," + B::sig() + ")";
  }
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
    if constexpr (std::is_same_v<typename X::T, double>) return gte_f64(x, Lo::eval(r, c, s)) && lte_f64(x, Hi::eval(r, c, s));
    else return x >= Lo::eval(r, c, s) && x <= Hi::eval(r, c, s);
  }
  static std::string sig() { return "between(" + X::sig() + "," + Lo::sig() + "," + Hi::sig() + ")"; }
};

template <class... As>
struct And {
  using T = bool;
  static constexpr uint32_t fmask = (As::fmask | ...), imask = (As::imask | ...), lmask = (As::lmask | ...);
  static constexpr bool uses_join = (As::uses_join || ...);
  // Null-free inputs: three-valued logic degenerates to &&. All conjuncts are evaluated (no
  // divergence); none of them can raise.
  __device__ static __forceinline__ bool eval(const PairRegs& r, const Consts& c, int s) { return (As::eval(r, c, s) & ...); }
  static std::string sig() {
    std::string out = "and(";
    bool first = true;
    ((out += (first ? "" : ",") + As::sig(), first = false), ...);
    return out + ")";
  }
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
  static std::string sig() { return "switch(" + C::sig() + "," + A::sig() + "," + B::sig() + ")"; }
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
  static_assert(((std::is_same_v<typename Ps::T, double>) && ...), "fused projections are DOUBLE");
  using F = Filter;
  template <int I>
  __device__ static __forceinline__ void project(const PairRegs& r, const Consts& c, int s, double (&out)[kNP]) {
    if constexpr (I < kNP) {
      using P = std::tuple_element_t<I, std::tuple<Ps...>>;
      out[I] = P::eval(r, c, s);
      project<I + 1>(r, c, s, out);
    }
  }
  static std::string sig() {
    std::string out = "F:" + Filter::sig() + ";P:";
    bool first = true;
    ((out += (first ? "" : "|") + Ps::sig(), first = false), ...);
    if (kJoin) out += ";J:l" + std::to_string(JoinCol);
    return out;
  }
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
  const int32_t* join_head;
  const int32_t* join_codes;
  const uint8_t* join_flag;
  int64_t join_min, join_range;
};

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
  // Predicated adds keep every accumulator in a register (no dynamic indexing).
  __device__ __forceinline__ void add(int gid, const double (&v)[P::kNP]) {
#pragma unroll
    for (int g = 0; g < kMaxG; ++g) {
      if (gid == g) {
        cnt[g] += 1;
#pragma unroll
        for (int p = 0; p < P::kNP; ++p) sum[g][p] = __dadd_rn(sum[g][p], v[p]);
      }
    }
  }
};

template <class P, int kMaxG, bool kPair, class KeyT>
__device__ __forceinline__ void process(const KernelArgs& a, const PairRegs& r0, const KeyT (&kv)[VB2_FUSED_MAX_KEYS][2],
                                        Accum<P, kMaxG>& acc) {
  PairRegs r = r0;
#pragma unroll
  for (int s = 0; s < (kPair ? 2 : 1); ++s) {
    bool keep = P::F::eval(r, a.consts, s);
    if constexpr (P::kJoin) {
      r.join_flag[s] = false;
      if (keep) {
        // Array-mode probe: slot = key - min; head holds build row + 1 (0 = no match).
        const int64_t key = r.l[P::kJoinCol][s];
        const int64_t slot = key - a.join_min;
        int32_t hit = 0;
        if (slot >= 0 && slot < a.join_range) hit = __ldg(a.join_head + slot);
        keep = hit != 0;
        if (keep) {
          const int32_t code = a.join_codes ? __ldg(a.join_codes + (hit - 1)) : (hit - 1);
          r.join_flag[s] = __ldg(a.join_flag + code) != 0;
        }
      }
    }
    int gid = -1;
    double v[P::kNP];
    if (keep) {
      gid = (kMaxG == 1) ? 0 : group_of(a, kv, s);
      P::template project<0>(r, a.consts, s, v);
    } else {
#pragma unroll
      for (int p = 0; p < P::kNP; ++p) v[p] = 0.0;
    }
    acc.add(gid, v);
  }
}

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
  // block reduction: shuffle within warps, then across warps in fixed order
  constexpr int kVals = kMaxG * (P::kNP + 1);
  __shared__ double smem[kThreads / kWarp][kVals];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int g = 0; g < kMaxG; ++g) {
#pragma unroll
    for (int p = 0; p < P::kNP; ++p) {
      double v = warp_sum(acc.sum[g][p]);
      if (lane == 0) smem[warp][g * (P::kNP + 1) + p] = v;
    }
    int64_t c = warp_sum(static_cast<int64_t>(acc.cnt[g]));
    if (lane == 0) smem[warp][g * (P::kNP + 1) + P::kNP] = __longlong_as_double(c);
  }
  __syncthreads();
  if (threadIdx.x < kVals) {
    const bool is_cnt = (threadIdx.x % (P::kNP + 1)) == P::kNP;
    double out;
    if (is_cnt) {
      int64_t c = 0;
      for (int w = 0; w < kThreads / kWarp; ++w) c += __double_as_longlong(smem[w][threadIdx.x]);
      out = __longlong_as_double(c);
    } else {
      double s = 0.0;
      for (int w = 0; w < kThreads / kWarp; ++w) s = __dadd_rn(s, smem[w][threadIdx.x]);
      out = s;
    }
    partials[static_cast<int64_t>(blockIdx.x) * kVals + threadIdx.x] = out;
  }
}

// Folds per-block partials in block order into the persistent accumulators.
__global__ void fused_finalize_kernel(const double* __restrict__ partials, int nblocks, int kvals, int np, int maxg,
                                      int ngroups, double* __restrict__ sums, int64_t* __restrict__ counts);

using LaunchFn = int (*)(const KernelArgs&, double* sums, int64_t* counts, void* ws, size_t ws_bytes, cudaStream_t);

struct Entry {
  std::string signature;
  int nproj;
  bool join;
  LaunchFn launch;
};

int register_pipeline(const Entry& e);

}  // namespace fx
}  // namespace vb2
