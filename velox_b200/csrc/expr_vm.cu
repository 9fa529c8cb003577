// Expression VM: one kernel evaluates a whole compiled ExprSet (filter and/or projections) per row.
//
// General path of the expression engine (any supported expression, any encoding, nulls, errors);
// the ahead-of-time fused pipelines in fused_scan.cu are the specialised path for null-free
// flat inputs feeding an aggregation. Both run on the device — there is no CPU evaluation.
//
// Semantics restated from the reference (paths into /root/reference/velox):
//   * default-null functions: result is NULL when any argument is NULL (expression/Expr.cpp:1235-1268)
//   * AND / OR: SQL three-valued, FALSE (TRUE) dominates NULL and errors (expression/ConjunctExpr.cpp:93-179)
//   * CASE / IF: a NULL condition counts as FALSE; only the chosen branch's value, null and error
//     matter (expression/SwitchExpr.cpp:101-152)
//   * integer arithmetic is checked (common/base/CheckedArithmetic.h:27-60); DOUBLE arithmetic is
//     plain IEEE-754 with one rounding per operation (functions/prestosql/Arithmetic.h:52-141)
//   * comparisons treat NaN as the largest value (type/FloatingPointUtil.h:52-98)
//   * a filter keeps a row only if the predicate is TRUE and not NULL (exec/OperatorUtils.cpp:238-248)
// Errors are carried as a poison bit per register and raised only if they reach a live output
// (row kept by the filter / projected), which is how the reference's per-row error capture
// inside AND / CASE behaves (ConjunctExpr.cpp:98-99,167-168).
#include "common.cuh"

namespace vb2 {

constexpr int kVmMaxInstrs = 256;
constexpr int kVmMaxConsts = 32;
constexpr int kVmMaxCols = 32;
constexpr int kVmMaxOuts = 32;
constexpr int kVmMaxRegs = 64;

enum VmErr : int { kErrOverflow = 1, kErrDivZero = 2, kErrCast = 3 };

struct VmArgs {
  vb2_instr instrs[kVmMaxInstrs];
  vb2_const consts[kVmMaxConsts];
  vb2_column cols[kVmMaxCols];
  vb2_output outs[kVmMaxOuts];
  int32_t n_instrs, n_filter_instrs, filter_reg, n_outs;
  int64_t n;            // rows (filter pass) or output rows (project pass)
  const int32_t* sel;   // project pass: input row of output k (NULL = identity)
  uint32_t* sel_bits;   // filter pass output
  int32_t* error_flag;
};

__device__ __forceinline__ bool decode(const vb2_column& c, int64_t row, int64_t& base) {
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

__device__ __forceinline__ uint64_t load_value(const vb2_column& c, int64_t base) {
  switch (c.type) {
    case VB2_BIGINT: return static_cast<uint64_t>(reinterpret_cast<const int64_t*>(c.values)[base]);
    case VB2_DOUBLE: return static_cast<uint64_t>(__double_as_longlong(reinterpret_cast<const double*>(c.values)[base]));
    case VB2_INTEGER: return static_cast<uint64_t>(static_cast<int64_t>(reinterpret_cast<const int32_t*>(c.values)[base]));
    case VB2_BOOLEAN: return bit_at(reinterpret_cast<const uint64_t*>(c.values), base) ? 1 : 0;
    default: return static_cast<uint64_t>(base);  // VARCHAR: registers never hold strings
  }
}

constexpr int kVmThreads = 256;

// Per-row interpreter state. The value registers live in shared memory, one column of
// kVmThreads words per VM register ([reg][thread]: dynamically indexed by the program, conflict
// free, and no local-memory traffic); the null / error poison masks stay in real registers.
struct RowState {
  uint64_t* base;  // &smem[threadIdx.x]
  uint64_t nullmask, errmask;
  int errcode;
  __device__ __forceinline__ uint64_t& r(int i) { return base[i * kVmThreads]; }
};

__device__ __forceinline__ double as_f64(uint64_t v) { return __longlong_as_double(static_cast<int64_t>(v)); }
__device__ __forceinline__ uint64_t from_f64(double d) { return static_cast<uint64_t>(__double_as_longlong(d)); }

// kPlain: the host has proved that no register of this program can be NULL or poisoned for this
// batch (no nullable input column, no NULL constant, no operation that can raise or produce NULL):
// the mask bookkeeping — most of the interpreter's instructions per operation — compiles away.
template <bool kPlain>
__device__ void run_program(const VmArgs& a, int n_instrs, int64_t row, RowState& st) {
  st.nullmask = 0;
  st.errmask = 0;
  st.errcode = 0;
  for (int pc = 0; pc < n_instrs; ++pc) {
    const vb2_instr in = a.instrs[pc];
    const uint64_t dbit = 1ull << in.dst;
    auto is_null = [&](int r) { return kPlain ? false : static_cast<bool>((st.nullmask >> r) & 1ull); };
    auto is_err = [&](int r) { return kPlain ? false : static_cast<bool>((st.errmask >> r) & 1ull); };
    bool rnull = false, rerr = false;
    uint64_t rv = 0;
    auto raise = [&](int code) { rerr = true; if (!st.errcode) st.errcode = code; };
    switch (in.op) {
      case VB2_OP_LOAD: {
        int64_t base;
        rnull = decode(a.cols[in.a], row, base);
        if (kPlain) rnull = false;
        rv = rnull ? 0 : load_value(a.cols[in.a], base);
        break;
      }
      case VB2_OP_CONST: {
        const vb2_const& c = a.consts[in.a];
        rnull = c.is_null != 0;
        rv = c.type == VB2_DOUBLE ? from_f64(c.d) : static_cast<uint64_t>(c.i);
        break;
      }
      case VB2_OP_NULL: rnull = true; break;
      case VB2_OP_ADD: case VB2_OP_SUB: case VB2_OP_MUL: case VB2_OP_DIV: case VB2_OP_MOD: {
        rnull = is_null(in.a) | is_null(in.b);
        rerr = is_err(in.a) | is_err(in.b);
        if (rnull || rerr) break;
        if (in.type == VB2_DOUBLE) {
          const double x = as_f64(st.r(in.a)), y = as_f64(st.r(in.b));
          double r;
          switch (in.op) {
            case VB2_OP_ADD: r = __dadd_rn(x, y); break;
            case VB2_OP_SUB: r = __dsub_rn(x, y); break;
            case VB2_OP_MUL: r = __dmul_rn(x, y); break;
            case VB2_OP_DIV: r = __ddiv_rn(x, y); break;
            default: r = fmod(x, y);
          }
          rv = from_f64(r);
        } else {
          const int64_t x = static_cast<int64_t>(st.r(in.a)), y = static_cast<int64_t>(st.r(in.b));
          int64_t r = 0;
          bool ovf = false;
          switch (in.op) {
            case VB2_OP_ADD: ovf = add_overflow_i64(x, y, &r); break;
            case VB2_OP_SUB: ovf = sub_overflow_i64(x, y, &r); break;
            case VB2_OP_MUL: ovf = mul_overflow_i64(x, y, &r); break;
            case VB2_OP_DIV:
              if (y == 0) { raise(kErrDivZero); }
              else if (x == INT64_MIN && y == -1) ovf = true;
              else r = x / y;
              break;
            default:
              if (y == 0) { raise(kErrDivZero); }
              else r = (y == -1) ? 0 : x % y;
          }
          if (in.type == VB2_INTEGER && !ovf && (r < INT32_MIN || r > INT32_MAX)) ovf = true;
          if (ovf) raise(kErrOverflow);
          rv = static_cast<uint64_t>(r);
        }
        break;
      }
      case VB2_OP_NEG: {
        rnull = is_null(in.a);
        rerr = is_err(in.a);
        if (rnull || rerr) break;
        if (in.type == VB2_DOUBLE) rv = from_f64(-as_f64(st.r(in.a)));
        else {
          const int64_t x = static_cast<int64_t>(st.r(in.a));
          const int64_t lo = in.type == VB2_INTEGER ? INT32_MIN : INT64_MIN;
          if (x == lo) raise(kErrOverflow);
          else rv = static_cast<uint64_t>(-x);
        }
        break;
      }
      case VB2_OP_LT: case VB2_OP_LTE: case VB2_OP_GT: case VB2_OP_GTE: case VB2_OP_EQ: case VB2_OP_NEQ: {
        rnull = is_null(in.a) | is_null(in.b);
        rerr = is_err(in.a) | is_err(in.b);
        if (rnull || rerr) break;
        const int op = in.op - VB2_OP_LT;
        if (in.type == VB2_DOUBLE) rv = cmp_f64(op, as_f64(st.r(in.a)), as_f64(st.r(in.b)));
        else rv = cmp_int<int64_t>(op, static_cast<int64_t>(st.r(in.a)), static_cast<int64_t>(st.r(in.b)));
        break;
      }
      case VB2_OP_BETWEEN: {
        rnull = is_null(in.a) | is_null(in.b) | is_null(in.c);
        rerr = is_err(in.a) | is_err(in.b) | is_err(in.c);
        if (rnull || rerr) break;
        if (in.type == VB2_DOUBLE) {
          const double x = as_f64(st.r(in.a));
          rv = gte_f64(x, as_f64(st.r(in.b))) && lte_f64(x, as_f64(st.r(in.c)));
        } else {
          const int64_t x = static_cast<int64_t>(st.r(in.a));
          rv = x >= static_cast<int64_t>(st.r(in.b)) && x <= static_cast<int64_t>(st.r(in.c));
        }
        break;
      }
      case VB2_OP_AND: case VB2_OP_OR: {
        const bool dominant = in.op == VB2_OP_OR;  // value that decides the result
        const bool an = is_null(in.a), bn = is_null(in.b), ae = is_err(in.a), be = is_err(in.b);
        const bool av = st.r(in.a) != 0, bv = st.r(in.b) != 0;
        const bool a_decides = !an && !ae && av == dominant;
        const bool b_decides = !bn && !be && bv == dominant;
        if (a_decides || b_decides) rv = dominant;
        else if (ae || be) rerr = true;
        else if (an || bn) rnull = true;
        else rv = !dominant;
        break;
      }
      case VB2_OP_NOT:
        rnull = is_null(in.a);
        rerr = is_err(in.a);
        rv = st.r(in.a) == 0;
        break;
      case VB2_OP_IS_NULL:
        rerr = is_err(in.a);
        rv = is_null(in.a);
        break;
      case VB2_OP_SELECT: {
        if (is_err(in.a)) { rerr = true; break; }
        const bool take = !is_null(in.a) && st.r(in.a) != 0;
        const int src = take ? in.b : in.c;
        if (src < 0) { rnull = true; break; }  // CASE without ELSE
        rnull = is_null(src);
        rerr = is_err(src);
        rv = st.r(src);
        break;
      }
      case VB2_OP_CAST: {
        rnull = is_null(in.a);
        rerr = is_err(in.a);
        if (rnull || rerr) break;
        const int from = in.b, to = in.type;
        const uint64_t v = st.r(in.a);
        if (from == to) rv = v;
        else if (to == VB2_BOOLEAN) rv = from == VB2_DOUBLE ? (as_f64(v) != 0.0) : (v != 0);  // folly::to<bool>: value != 0 (NaN -> true)
        else if (to == VB2_DOUBLE) rv = from_f64(static_cast<double>(static_cast<int64_t>(v)));
        else if (from == VB2_DOUBLE) {
          const double d = as_f64(v);
          if (isnan(d)) { raise(kErrCast); break; }
          const double r = round(d);
          const double lo = to == VB2_INTEGER ? -2147483648.0 : -9223372036854775808.0;
          if (r < lo || r >= -lo) { raise(kErrCast); break; }
          rv = static_cast<uint64_t>(static_cast<int64_t>(r));
        } else if (to == VB2_INTEGER) {
          const int64_t x = static_cast<int64_t>(v);
          if (x < INT32_MIN || x > INT32_MAX) { raise(kErrCast); break; }
          rv = v;
        } else rv = v;  // INTEGER/BOOLEAN -> BIGINT (already sign-extended)
        break;
      }
      case VB2_OP_LIKE: case VB2_OP_STRCMP: {
        const vb2_column& c = a.cols[in.a];
        const vb2_const& k = a.consts[in.b];
        int64_t base;
        rnull = decode(c, row, base) || k.is_null;
        if (rnull) break;
        const int32_t* off = reinterpret_cast<const int32_t*>(c.values);
        const char* s = reinterpret_cast<const char*>(c.aux) + off[base];
        const int sl = off[base + 1] - off[base];
        if (in.op == VB2_OP_LIKE) rv = like_match(s, sl, k.str, k.len);
        else rv = cmp_int<int>(in.c, str_compare(s, sl, k.str, k.len), 0);
        break;
      }
      default: break;
    }
    st.r(in.dst) = rv;
    if (!kPlain) {
      st.nullmask = rnull ? (st.nullmask | dbit) : (st.nullmask & ~dbit);
      st.errmask = rerr ? (st.errmask | dbit) : (st.errmask & ~dbit);
    }
  }
}

__device__ __forceinline__ int user_code(int vmerr) { return vmerr ? vmerr : kErrOverflow; }

template <bool kPlain>
__global__ void __launch_bounds__(kVmThreads) vm_filter_kernel(const __grid_constant__ VmArgs a) {
  // each warp produces one 32-bit word of the selection bitmap per iteration
  const int64_t nwords = (a.n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * kVmThreads + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * kVmThreads) >> 5;
  extern __shared__ uint64_t vm_regs[];
  RowState st;
  st.base = vm_regs + threadIdx.x;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t row = (w << 5) + lane;
    bool keep = false;
    if (row < a.n) {
      run_program<kPlain>(a, a.n_filter_instrs, row, st);
      const bool err = (st.errmask >> a.filter_reg) & 1ull;
      const bool null = (st.nullmask >> a.filter_reg) & 1ull;
      if (err) atomicCAS(a.error_flag, 0, user_code(st.errcode));
      keep = !err && !null && st.r(a.filter_reg) != 0;
    }
    const unsigned word = __ballot_sync(0xffffffffu, keep);
    if (lane == 0) a.sel_bits[w] = word;
  }
}

template <bool kPlain>
__global__ void __launch_bounds__(kVmThreads) vm_project_kernel(const __grid_constant__ VmArgs a) {
  const int64_t nwords = (a.n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * kVmThreads + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * kVmThreads) >> 5;
  extern __shared__ uint64_t vm_regs[];
  RowState st;
  st.base = vm_regs + threadIdx.x;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t k = (w << 5) + lane;
    const bool live = k < a.n;
    if (live) {
      const int64_t row = a.sel ? a.sel[k] : k;
      run_program<kPlain>(a, a.n_instrs, row, st);
    }
    for (int o = 0; o < a.n_outs; ++o) {
      const vb2_output& out = a.outs[o];
      bool valid = false;
      if (live) {
        const bool err = (st.errmask >> out.reg) & 1ull;
        if (err) atomicCAS(a.error_flag, 0, user_code(st.errcode));
        valid = !err && !((st.nullmask >> out.reg) & 1ull);
        const uint64_t v = valid ? st.r(out.reg) : 0;
        switch (out.type) {
          case VB2_INTEGER: reinterpret_cast<int32_t*>(out.values)[k] = static_cast<int32_t>(v); break;
          case VB2_BOOLEAN: reinterpret_cast<uint8_t*>(out.values)[k] = static_cast<uint8_t>(v); break;
          default: reinterpret_cast<uint64_t*>(out.values)[k] = v;
        }
      }
      const unsigned word = __ballot_sync(0xffffffffu, valid);
      if (lane == 0) reinterpret_cast<uint32_t*>(out.nulls)[w] = word;
    }
  }
}

// ---- selection bitmap -> ascending row numbers ------------------------------------------------
constexpr int kSelThreads = 256;
constexpr int kSelWordsPerBlock = 1024;  // 32-bit words -> 32768 rows per block

__global__ void sel_count_kernel(const uint32_t* __restrict__ bits, int64_t nwords, int32_t* __restrict__ block_counts) {
  __shared__ int warp_sums[kSelThreads / kWarp];
  const int64_t w0 = static_cast<int64_t>(blockIdx.x) * kSelWordsPerBlock;
  int c = 0;
  for (int i = threadIdx.x; i < kSelWordsPerBlock; i += kSelThreads)
    if (w0 + i < nwords) c += __popc(bits[w0 + i]);
  c = warp_sum(c);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = c;
  __syncthreads();
  if (threadIdx.x == 0) {
    int s = 0;
    for (int w = 0; w < kSelThreads / kWarp; ++w) s += warp_sums[w];
    block_counts[blockIdx.x] = s;
  }
}
// single block exclusive scan over block counts (nblocks = rows / 32768: small)
__global__ void sel_scan_kernel(const int32_t* __restrict__ block_counts, int64_t nblocks, int64_t* __restrict__ block_offsets,
                                int64_t* __restrict__ total) {
  __shared__ int64_t carry;
  __shared__ int64_t tmp[1024];
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (int64_t b0 = 0; b0 < nblocks; b0 += 1024) {
    const int64_t b = b0 + threadIdx.x;
    const int64_t v = b < nblocks ? block_counts[b] : 0;
    tmp[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {
      const int64_t t = threadIdx.x >= o ? tmp[threadIdx.x - o] : 0;
      __syncthreads();
      tmp[threadIdx.x] += t;
      __syncthreads();
    }
    if (b < nblocks) block_offsets[b] = carry + tmp[threadIdx.x] - v;
    __syncthreads();
    if (threadIdx.x == 1023) carry += tmp[1023];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = carry;
}
__global__ void sel_write_kernel(const uint32_t* __restrict__ bits, int64_t nwords, const int64_t* __restrict__ block_offsets,
                                 int32_t* __restrict__ indices) {
  __shared__ int word_prefix[kSelWordsPerBlock];
  __shared__ int warp_tot[kSelThreads / kWarp];
  const int64_t w0 = static_cast<int64_t>(blockIdx.x) * kSelWordsPerBlock;
  // exclusive prefix of popcounts over the block's words: 4 consecutive words per thread
  const int t = threadIdx.x;
  int pc[4], run = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int64_t w = w0 + t * 4 + j;
    pc[j] = w < nwords ? __popc(bits[w]) : 0;
    run += pc[j];
  }
  int incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if ((t & 31) >= o) incl += v;
  }
  if ((t & 31) == 31) warp_tot[t >> 5] = incl;
  __syncthreads();
  int base = 0;
  for (int w = 0; w < (t >> 5); ++w) base += warp_tot[w];
  int excl = base + incl - run;
#pragma unroll
  for (int j = 0; j < 4; ++j) { word_prefix[t * 4 + j] = excl; excl += pc[j]; }
  __syncthreads();
  const int64_t out0 = block_offsets[blockIdx.x];
  // Sparse blocks (selective filters: most words are zero): every thread expands its own four
  // words bit by bit — few bits, so the scattered stores do not matter and no warp walks 128 empty words.
  __shared__ int block_total;
  if (t == kSelThreads - 1) block_total = excl;
  __syncthreads();
  if (block_total < kSelWordsPerBlock * 4) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int64_t w = w0 + t * 4 + j;
      if (w >= nwords) break;
      uint32_t word = bits[w];
      int64_t pos = out0 + word_prefix[t * 4 + j];
      while (word) {
        const int b = __ffs(word) - 1;
        indices[pos++] = static_cast<int32_t>((w << 5) + b);
        word &= word - 1;
      }
    }
    return;
  }
  // One 32-row word per warp step: lane b owns row (w << 5) + b and stores it at its rank among
  // the set bits, so a warp's stores are consecutive (a thread-per-word bit loop wrote 32
  // scattered sequences per warp and ran at a fifth of the bandwidth).
  const int lane = t & 31, warp = t >> 5;
  for (int i = warp; i < kSelWordsPerBlock; i += kSelThreads / kWarp) {
    const int64_t w = w0 + i;
    if (w >= nwords) break;
    const uint32_t word = bits[w];
    if ((word >> lane) & 1u) indices[out0 + word_prefix[i] + __popc(word & ((1u << lane) - 1u))] = static_cast<int32_t>((w << 5) + lane);
  }
}

static unsigned vm_grid(int64_t n) {
  int64_t b = (n + kVmThreads - 1) / kVmThreads;
  int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

// Dynamic shared memory of a VM launch: one kVmThreads-word column per VM register.
template <class K>
static int vm_smem(K kernel, const vb2_program* prog, size_t* bytes) {
  int regs = 1;
  for (int i = 0; i < prog->n_instrs; ++i) {
    const vb2_instr& in = prog->instrs[i];
    regs = std::max(regs, std::max(std::max(in.dst, in.a), std::max(in.b, in.c)) + 1);  // operand fields are < 64 whatever they mean
  }
  regs = std::min(std::max(regs, static_cast<int>(prog->n_regs)), kVmMaxRegs);
  *bytes = static_cast<size_t>(regs) * kVmThreads * sizeof(uint64_t);
  if (*bytes > 48 * 1024) {
    static size_t configured = 0;
    if (configured < *bytes) {
      VB2_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kVmMaxRegs * kVmThreads * sizeof(uint64_t))));
      configured = kVmMaxRegs * kVmThreads * sizeof(uint64_t);
    }
  }
  return VB2_OK;
}

// True when no register can become NULL or poisoned: every loaded column is free of NULLs, no
// NULL constant, and only operations that neither raise nor produce NULL from non-NULL inputs
// (double arithmetic except nothing, comparisons, BETWEEN, AND / OR / NOT, IS_NULL, CASE with ELSE,
// LIKE / string compare against non-NULL patterns).
// Registered device functions exist as source text only: the interpreter cannot run them.
static bool calls_registered_function(const vb2_program* prog, int n_instrs) {
  for (int i = 0; i < n_instrs; ++i)
    if (prog->instrs[i].op == VB2_OP_CALL) return true;
  return false;
}
static const char kNeedsJit[] = "the expression calls a registered device function, which needs the expression JIT (NVRTC missing, disabled, or the function's source does not compile)";

static bool plain_program(const vb2_program* prog, int n_instrs, const vb2_column* cols) {
  for (int i = 0; i < n_instrs; ++i) {
    const vb2_instr& in = prog->instrs[i];
    switch (in.op) {
      case VB2_OP_LOAD: case VB2_OP_LIKE: case VB2_OP_STRCMP: {
        const vb2_column& c = cols[in.a];
        if (c.nulls || c.dict_nulls) return false;
        if (in.op != VB2_OP_LOAD && prog->consts[in.b].is_null) return false;
        break;
      }
      case VB2_OP_CONST:
        if (prog->consts[in.a].is_null) return false;
        break;
      case VB2_OP_ADD: case VB2_OP_SUB: case VB2_OP_MUL: case VB2_OP_DIV: case VB2_OP_MOD: case VB2_OP_NEG:
        if (in.type != VB2_DOUBLE) return false;  // checked integer arithmetic can raise
        break;
      case VB2_OP_LT: case VB2_OP_LTE: case VB2_OP_GT: case VB2_OP_GTE: case VB2_OP_EQ: case VB2_OP_NEQ: case VB2_OP_BETWEEN:
      case VB2_OP_AND: case VB2_OP_OR: case VB2_OP_NOT: case VB2_OP_IS_NULL:
        break;
      case VB2_OP_SELECT:
        if (in.c < 0) return false;  // CASE without ELSE yields NULL
        break;
      case VB2_OP_CAST:
        if (!(in.type == VB2_DOUBLE || in.type == VB2_BOOLEAN || in.b == in.type || (in.type == VB2_BIGINT && in.b != VB2_DOUBLE))) return false;  // narrowing casts can raise
        break;
      default:
        return false;
    }
  }
  return true;
}

namespace jit {
// expr_jit.cu: VB2_ERR_UNSUPPORTED = run the interpreter
int launch(const vb2_program* p, const vb2_column* cols, int ncols, bool filter, const int32_t* sel, int64_t n, uint32_t* sel_bits,
           const vb2_output* outs, int nouts, int32_t* error_flag, cudaStream_t st);
}  // namespace jit

static int fill_args(VmArgs& a, const vb2_program* prog, const vb2_column* cols, int32_t ncols) {
  if (!prog || prog->n_instrs < 0 || prog->n_instrs > kVmMaxInstrs) return fail_msg(VB2_ERR_UNSUPPORTED, "expression program too long (max 256 instructions)");
  if (prog->n_consts > kVmMaxConsts) return fail_msg(VB2_ERR_UNSUPPORTED, "too many constants (max 32)");
  if (ncols > kVmMaxCols) return fail_msg(VB2_ERR_UNSUPPORTED, "too many input columns (max 32)");
  if (prog->n_regs > kVmMaxRegs) return fail_msg(VB2_ERR_UNSUPPORTED, "too many registers (max 64)");
  for (int i = 0; i < prog->n_instrs; ++i) a.instrs[i] = prog->instrs[i];
  for (int i = 0; i < prog->n_consts; ++i) a.consts[i] = prog->consts[i];
  for (int i = 0; i < ncols; ++i) a.cols[i] = cols[i];
  a.n_instrs = prog->n_instrs;
  a.n_filter_instrs = prog->n_filter_instrs;
  a.filter_reg = prog->filter_reg;
  return VB2_OK;
}

}  // namespace vb2

using namespace vb2;

extern "C" {

int vb2k_eval_filter(const vb2_program* prog, const vb2_column* cols, int32_t ncols, int64_t rows,
                     uint64_t* sel_bits, int32_t* error_flag, void* stream) {
  if (rows <= 0) return VB2_OK;
  static thread_local VmArgs a;
  int rc = fill_args(a, prog, cols, ncols);
  if (rc) return rc;
  if (prog->filter_reg < 0) return fail_msg(VB2_ERR_INVALID, "eval_filter: program has no filter");
  a.n = rows;
  a.sel = nullptr;
  a.sel_bits = reinterpret_cast<uint32_t*>(sel_bits);
  a.error_flag = error_flag;
  a.n_outs = 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // zero the tail word so that bits beyond `rows` read as 0 for 64-bit consumers
  const int64_t nwords64 = (rows + 63) >> 6;
  VB2_CUDA_OK(cudaMemsetAsync(sel_bits + nwords64 - 1, 0, sizeof(uint64_t), st));
  rc = jit::launch(prog, cols, ncols, true, nullptr, rows, reinterpret_cast<uint32_t*>(sel_bits), nullptr, 0, error_flag, st);
  if (rc != VB2_ERR_UNSUPPORTED) return rc;
  if (calls_registered_function(prog, prog->n_filter_instrs)) return fail_msg(VB2_ERR_UNSUPPORTED, kNeedsJit);
  size_t smem = 0;
  if (plain_program(prog, prog->n_filter_instrs, cols)) {
    if ((rc = vm_smem(vm_filter_kernel<true>, prog, &smem))) return rc;
    vm_filter_kernel<true><<<vb2::counted(vm_grid(rows)), kVmThreads, smem, st>>>(a);
  } else {
    if ((rc = vm_smem(vm_filter_kernel<false>, prog, &smem))) return rc;
    vm_filter_kernel<false><<<vb2::counted(vm_grid(rows)), kVmThreads, smem, st>>>(a);
  }
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

size_t vb2k_bits_to_indices_workspace(int64_t rows) {
  const int64_t nwords = (rows + 31) >> 5;
  const int64_t nblocks = (nwords + kSelWordsPerBlock - 1) / kSelWordsPerBlock;
  return static_cast<size_t>(nblocks) * (sizeof(int32_t) + sizeof(int64_t)) + 64;
}

int vb2k_bits_to_indices(const uint64_t* sel_bits, int64_t rows, int32_t* indices, int64_t* count_out,
                         void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (rows <= 0) {
    VB2_CUDA_OK(cudaMemsetAsync(count_out, 0, sizeof(int64_t), st));
    return VB2_OK;
  }
  if (workspace_bytes < vb2k_bits_to_indices_workspace(rows)) return fail_msg(VB2_ERR_INVALID, "bits_to_indices: workspace too small");
  const int64_t nwords = (rows + 31) >> 5;
  const int64_t nblocks = (nwords + kSelWordsPerBlock - 1) / kSelWordsPerBlock;
  int64_t* offsets = reinterpret_cast<int64_t*>(workspace);
  int32_t* counts = reinterpret_cast<int32_t*>(offsets + nblocks);
  const uint32_t* bits = reinterpret_cast<const uint32_t*>(sel_bits);
  sel_count_kernel<<<vb2::counted(static_cast<unsigned>(nblocks)), kSelThreads, 0, st>>>(bits, nwords, counts);
  sel_scan_kernel<<<vb2::counted(1), 1024, 0, st>>>(counts, nblocks, offsets, count_out);
  sel_write_kernel<<<vb2::counted(static_cast<unsigned>(nblocks)), kSelThreads, 0, st>>>(bits, nwords, offsets, indices);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_eval_project(const vb2_program* prog, const vb2_column* cols, int32_t ncols, const int32_t* sel,
                      int64_t n, const vb2_output* outs, int32_t nouts, int32_t* error_flag, void* stream) {
  if (n <= 0) return VB2_OK;
  static thread_local VmArgs a;
  int rc = fill_args(a, prog, cols, ncols);
  if (rc) return rc;
  if (nouts > kVmMaxOuts) return fail_msg(VB2_ERR_UNSUPPORTED, "too many outputs (max 32)");
  for (int i = 0; i < nouts; ++i) a.outs[i] = outs[i];
  a.n_outs = nouts;
  a.n = n;
  a.sel = sel;
  a.sel_bits = nullptr;
  a.error_flag = error_flag;
  rc = jit::launch(prog, cols, ncols, false, sel, n, nullptr, outs, nouts, error_flag, static_cast<cudaStream_t>(stream));
  if (rc != VB2_ERR_UNSUPPORTED) return rc;
  if (calls_registered_function(prog, prog->n_instrs)) return fail_msg(VB2_ERR_UNSUPPORTED, kNeedsJit);
  size_t smem = 0;
  if (plain_program(prog, prog->n_instrs, cols)) {
    if ((rc = vm_smem(vm_project_kernel<true>, prog, &smem))) return rc;
    vm_project_kernel<true><<<vb2::counted(vm_grid(n)), kVmThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
  } else {
    if ((rc = vm_smem(vm_project_kernel<false>, prog, &smem))) return rc;
    vm_project_kernel<false><<<vb2::counted(vm_grid(n)), kVmThreads, smem, static_cast<cudaStream_t>(stream)>>>(a);
  }
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

}  // extern "C"
