// Expression JIT: a linear expression program (vb2_program) is turned into straight-line CUDA C++
// — one scalar triple (value, is-null, poisoned) per VM register, column decoding specialised on
// type / encoding / presence of validity bitmaps — compiled for sm_100a with NVRTC and cached.
//
// This is the device-side counterpart of ExprSet's compiled expression tree
// (velox/expression/Expr.cpp:2339 ExprSet::eval; the reference's own GPU prototype JIT-compiles
// too, velox/experimental/wave/jit). The interpreter in expr_vm.cu stays: it runs any program
// immediately and is the parity partner of the JIT in the tests; the JIT removes the per-operation
// decode, dynamic register indexing and mask bookkeeping, which is what keeps the interpreter
// issue-bound instead of HBM-bound. Semantics come from the same source text as the interpreter:
// vm_ops.inc is compiled into both.
//
// NVRTC is loaded lazily with dlopen (the library must load on hosts without a GPU); kernels are
// loaded with the runtime API's cudaLibraryLoadData. Any failure to JIT (NVRTC missing, an
// operation the generator does not cover) falls back to the interpreter kernels — still CUDA.

#include <cstdio>
#include <cstdlib>
#include <memory>
#include <mutex>
#include <sstream>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"
#include "jit_common.h"
#include "vm_ops_str.h"  // generated from vm_ops.inc: static const char kVmOpsSource[]

namespace vb2 {
namespace jit {

constexpr int kMaxCols = 32, kMaxConsts = 32, kMaxOuts = 32, kThreads = 256;

// Kernel argument block (same text in the generated source).
struct JitArgs {
  vb2_column cols[kMaxCols];
  vb2_const consts[kMaxConsts];
  vb2_output outs[kMaxOuts];
  long long n;
  const int* sel;
  unsigned* sel_bits;
  int* error_flag;
};

static const char kPrelude[] = R"SRC(
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef unsigned char uint8_t;
#define INT64_MIN (-9223372036854775807LL - 1)
#define INT32_MIN (-2147483647 - 1)
#define INT32_MAX 2147483647
#include "vm_ops.inc"
struct vb2_column { int32_t type; int32_t encoding; int64_t size; const void* values; const uint64_t* nulls; const int32_t* indices;
                    int64_t dict_size; const uint64_t* dict_nulls; const void* aux; };
struct vb2_const { int32_t type; int32_t is_null; int64_t i; double d; const char* str; int32_t len; int32_t pad; };
struct vb2_output { int32_t reg; int32_t type; void* values; uint64_t* nulls; };
struct JitArgs { vb2_column cols[32]; vb2_const consts[32]; vb2_output outs[32]; long long n; const int* sel; unsigned* sel_bits; int* error_flag; };
static_assert(sizeof(vb2_column) == VB2_SIZEOF_COLUMN && sizeof(vb2_const) == VB2_SIZEOF_CONST && sizeof(vb2_output) == VB2_SIZEOF_OUTPUT &&
              sizeof(JitArgs) == VB2_SIZEOF_ARGS, "argument block layout differs between host and JIT");
__device__ __forceinline__ double as_f64(uint64_t v) { return __longlong_as_double((long long)v); }
__device__ __forceinline__ uint64_t from_f64(double d) { return (uint64_t)__double_as_longlong(d); }
)SRC";

// ---- registered device functions (VB2_OP_CALL) -------------------------------------------------
struct DeviceFn {
  std::string entry, source;
  int ret = 0, nargs = 0;
  int args[3] = {0, 0, 0};
};
static std::mutex g_fn_mu;
static std::vector<DeviceFn> g_fns;

static bool device_fn(int id, DeviceFn* out) {
  std::lock_guard<std::mutex> lock(g_fn_mu);
  if (id < 0 || id >= static_cast<int>(g_fns.size())) return false;
  *out = g_fns[id];
  return true;
}
static const char* ctype_of(int t) {
  switch (t) {
    case VB2_BIGINT: return "long long";
    case VB2_INTEGER: return "int";
    case VB2_DOUBLE: return "double";
    case VB2_BOOLEAN: return "bool";
    default: return nullptr;
  }
}

// ---- code generation -----------------------------------------------------------------------------
struct Gen {
  std::ostringstream o;
  const vb2_program* prog;
  const vb2_column* cols;
  bool ok = true;
  std::vector<int> used_fns;  // registered functions this kernel calls (their source goes in front)

  static std::string V(int r) { return "v" + std::to_string(r); }
  static std::string N(int r) { return "n" + std::to_string(r); }
  static std::string E(int r) { return "e" + std::to_string(r); }
  std::string C(int c) const { return "a.cols[" + std::to_string(c) + "]"; }
  std::string K(int k) const { return "a.consts[" + std::to_string(k) + "]"; }

  // value expression of a column element at `b`
  std::string loadv(int c, const std::string& b) const {
    const std::string vals = C(c) + ".values";
    switch (cols[c].type) {
      case VB2_BIGINT: return "(uint64_t)((const long long*)" + vals + ")[" + b + "]";
      case VB2_DOUBLE: return "from_f64(((const double*)" + vals + ")[" + b + "])";
      case VB2_INTEGER: return "(uint64_t)(long long)((const int*)" + vals + ")[" + b + "]";
      case VB2_BOOLEAN: return "(uint64_t)(bit_at((const uint64_t*)" + vals + ", " + b + ") ? 1 : 0)";
      default: return "(uint64_t)" + b;  // VARCHAR: registers never hold strings
    }
  }
  // emits `long long b; bool rn;` = decoded base index and null flag of column c at `row`
  void decode(int c) {
    const vb2_column& col = cols[c];
    o << "      long long b; bool rn;\n";
    if (col.encoding == VB2_FLAT) {
      o << "      b = row; rn = " << (col.nulls ? "!bit_at(" + C(c) + ".nulls, row)" : std::string("false")) << ";\n";
    } else if (col.encoding == VB2_DICTIONARY) {
      o << "      rn = " << (col.nulls ? "!bit_at(" + C(c) + ".nulls, row)" : std::string("false")) << "; b = 0;\n";
      o << "      if (!rn) { b = " << C(c) << ".indices[row];";
      if (col.dict_nulls) o << " rn = !bit_at(" << C(c) << ".dict_nulls, b);";
      o << " }\n";
    } else {
      o << "      b = 0; rn = " << (col.nulls ? "!bit_at(" + C(c) + ".nulls, 0)" : std::string("false")) << ";\n";
    }
  }

  void instr(const vb2_instr& in) {
    const int d = in.dst, A = in.a, B = in.b, Cc = in.c;
    const bool dbl = in.type == VB2_DOUBLE;
    o << "    {  // op " << in.op << "\n      uint64_t rv = 0; bool rnull = false, rerr = false;\n";
    auto nulls2 = [&] { o << "      rnull = " << N(A) << " | " << N(B) << "; rerr = " << E(A) << " | " << E(B) << ";\n"; };
    auto raise = [](const char* code) { return std::string("{ rerr = true; if (!errcode) errcode = ") + code + "; }"; };
    switch (in.op) {
      case VB2_OP_LOAD:
        o << "      {\n";
        decode(A);
        o << "      rnull = rn; rv = rn ? 0 : " << loadv(A, "b") << ";\n      }\n";
        break;
      case VB2_OP_CONST:
        o << "      rnull = " << (prog->consts[A].is_null ? "true" : "false") << ";\n";
        o << "      rv = " << (prog->consts[A].type == VB2_DOUBLE ? "from_f64(" + K(A) + ".d)" : "(uint64_t)" + K(A) + ".i") << ";\n";
        break;
      case VB2_OP_NULL: o << "      rnull = true;\n"; break;
      case VB2_OP_ADD: case VB2_OP_SUB: case VB2_OP_MUL: case VB2_OP_DIV: case VB2_OP_MOD:
        nulls2();
        o << "      if (!(rnull || rerr)) {\n";
        if (dbl) {
          const char* f = in.op == VB2_OP_ADD ? "__dadd_rn(x, y)" : in.op == VB2_OP_SUB ? "__dsub_rn(x, y)" : in.op == VB2_OP_MUL ? "__dmul_rn(x, y)"
                          : in.op == VB2_OP_DIV ? "__ddiv_rn(x, y)" : "fmod(x, y)";
          o << "        const double x = as_f64(" << V(A) << "), y = as_f64(" << V(B) << ");\n        rv = from_f64(" << f << ");\n";
        } else {
          o << "        const long long x = (long long)" << V(A) << ", y = (long long)" << V(B) << "; long long r = 0; bool ovf = false;\n";
          switch (in.op) {
            case VB2_OP_ADD: o << "        ovf = add_overflow_i64(x, y, &r);\n"; break;
            case VB2_OP_SUB: o << "        ovf = sub_overflow_i64(x, y, &r);\n"; break;
            case VB2_OP_MUL: o << "        ovf = mul_overflow_i64(x, y, &r);\n"; break;
            case VB2_OP_DIV:
              o << "        if (y == 0) " << raise("2") << " else if (x == INT64_MIN && y == -1) ovf = true; else r = x / y;\n";
              break;
            default:
              o << "        if (y == 0) " << raise("2") << " else r = (y == -1) ? 0 : x % y;\n";
          }
          if (in.type == VB2_INTEGER) o << "        if (!ovf && (r < INT32_MIN || r > INT32_MAX)) ovf = true;\n";
          o << "        if (ovf) " << raise("1") << "\n        rv = (uint64_t)r;\n";
        }
        o << "      }\n";
        break;
      case VB2_OP_NEG:
        o << "      rnull = " << N(A) << "; rerr = " << E(A) << ";\n      if (!(rnull || rerr)) {\n";
        if (dbl) o << "        rv = from_f64(-as_f64(" << V(A) << "));\n";
        else
          o << "        const long long x = (long long)" << V(A) << ";\n        if (x == " << (in.type == VB2_INTEGER ? "(long long)INT32_MIN" : "INT64_MIN")
            << ") " << raise("1") << " else rv = (uint64_t)(-x);\n";
        o << "      }\n";
        break;
      case VB2_OP_LT: case VB2_OP_LTE: case VB2_OP_GT: case VB2_OP_GTE: case VB2_OP_EQ: case VB2_OP_NEQ:
        nulls2();
        o << "      if (!(rnull || rerr)) rv = ";
        if (dbl) o << "cmp_f64(" << (in.op - VB2_OP_LT) << ", as_f64(" << V(A) << "), as_f64(" << V(B) << "));\n";
        else o << "cmp_int<long long>(" << (in.op - VB2_OP_LT) << ", (long long)" << V(A) << ", (long long)" << V(B) << ");\n";
        break;
      case VB2_OP_BETWEEN:
        o << "      rnull = " << N(A) << " | " << N(B) << " | " << N(Cc) << "; rerr = " << E(A) << " | " << E(B) << " | " << E(Cc) << ";\n";
        o << "      if (!(rnull || rerr)) {\n";
        if (dbl) o << "        const double x = as_f64(" << V(A) << ");\n        rv = gte_f64(x, as_f64(" << V(B) << ")) && lte_f64(x, as_f64(" << V(Cc) << "));\n";
        else o << "        const long long x = (long long)" << V(A) << ";\n        rv = x >= (long long)" << V(B) << " && x <= (long long)" << V(Cc) << ";\n";
        o << "      }\n";
        break;
      case VB2_OP_AND: case VB2_OP_OR: {
        const char* dom = in.op == VB2_OP_OR ? "true" : "false";
        o << "      const bool dominant = " << dom << ";\n";
        o << "      const bool an = " << N(A) << ", bn = " << N(B) << ", ae = " << E(A) << ", be = " << E(B) << ";\n";
        o << "      const bool av = " << V(A) << " != 0, bv = " << V(B) << " != 0;\n";
        o << "      const bool a_decides = !an && !ae && av == dominant, b_decides = !bn && !be && bv == dominant;\n";
        o << "      if (a_decides || b_decides) rv = dominant; else if (ae || be) rerr = true; else if (an || bn) rnull = true; else rv = !dominant;\n";
        break;
      }
      case VB2_OP_NOT: o << "      rnull = " << N(A) << "; rerr = " << E(A) << "; rv = " << V(A) << " == 0;\n"; break;
      case VB2_OP_IS_NULL: o << "      rerr = " << E(A) << "; rv = " << N(A) << ";\n"; break;
      case VB2_OP_SELECT:
        o << "      if (" << E(A) << ") rerr = true; else {\n        const bool take = !" << N(A) << " && " << V(A) << " != 0;\n";
        if (Cc < 0) o << "        if (take) { rnull = " << N(B) << "; rerr = " << E(B) << "; rv = " << V(B) << "; } else rnull = true;\n";
        else
          o << "        rnull = take ? " << N(B) << " : " << N(Cc) << "; rerr = take ? " << E(B) << " : " << E(Cc) << "; rv = take ? " << V(B) << " : " << V(Cc)
            << ";\n";
        o << "      }\n";
        break;
      case VB2_OP_CAST: {
        const int from = B, to = in.type;
        o << "      rnull = " << N(A) << "; rerr = " << E(A) << ";\n      if (!(rnull || rerr)) {\n        const uint64_t v = " << V(A) << ";\n";
        if (from == to) o << "        rv = v;\n";
        else if (to == VB2_BOOLEAN) o << (from == VB2_DOUBLE ? "        rv = as_f64(v) != 0.0;\n" : "        rv = v != 0;\n");
        else if (to == VB2_DOUBLE) o << "        rv = from_f64((double)(long long)v);\n";
        else if (from == VB2_DOUBLE) {
          o << "        const double dd = as_f64(v);\n        if (isnan(dd)) " << raise("3") << " else {\n          const double r = round(dd);\n";
          o << "          const double lo = " << (to == VB2_INTEGER ? "-2147483648.0" : "-9223372036854775808.0") << ";\n";
          o << "          if (r < lo || r >= -lo) " << raise("3") << " else rv = (uint64_t)(long long)r;\n        }\n";
        } else if (to == VB2_INTEGER) {
          o << "        const long long x = (long long)v;\n        if (x < INT32_MIN || x > INT32_MAX) " << raise("3") << " else rv = v;\n";
        } else o << "        rv = v;\n";
        o << "      }\n";
        break;
      }
      case VB2_OP_LIKE: case VB2_OP_STRCMP:
        if (cols[A].type != VB2_VARCHAR) { ok = false; break; }
        o << "      {\n";
        decode(A);
        o << "      rnull = rn || " << (prog->consts[B].is_null ? "true" : "false") << ";\n";
        o << "      if (!rnull) {\n        const int* off = (const int*)" << C(A) << ".values;\n        const char* s = (const char*)" << C(A)
          << ".aux + off[b];\n        const int sl = off[b + 1] - off[b];\n";
        if (in.op == VB2_OP_LIKE) o << "        rv = like_match(s, sl, " << K(B) << ".str, " << K(B) << ".len);\n";
        else o << "        rv = cmp_int<int>(" << Cc << ", str_compare(s, sl, " << K(B) << ".str, " << K(B) << ".len), 0);\n";
        o << "      }\n      }\n";
        break;
      case VB2_OP_CALL: {
        DeviceFn f;
        if (!device_fn(VB2_CALL_FN(in.type), &f)) { ok = false; break; }
        const int regs[3] = {A, B, Cc};
        o << "      rnull = false; rerr = false;\n";
        for (int i = 0; i < f.nargs; ++i) o << "      rnull |= " << N(regs[i]) << "; rerr |= " << E(regs[i]) << ";\n";
        o << "      if (!(rnull || rerr)) {\n        const " << ctype_of(f.ret) << " r = " << f.entry << "(";
        for (int i = 0; i < f.nargs; ++i) {
          if (i) o << ", ";
          if (f.args[i] == VB2_DOUBLE) o << "as_f64(" << V(regs[i]) << ")";
          else if (f.args[i] == VB2_BOOLEAN) o << "(" << V(regs[i]) << " != 0)";
          else o << "(" << ctype_of(f.args[i]) << ")(long long)" << V(regs[i]);
        }
        o << ");\n        rv = " << (f.ret == VB2_DOUBLE ? "from_f64(r)" : f.ret == VB2_BOOLEAN ? "(uint64_t)(r ? 1 : 0)" : "(uint64_t)(long long)r") << ";\n      }\n";
        bool seen = false;
        for (int u : used_fns) seen = seen || u == VB2_CALL_FN(in.type);
        if (!seen) used_fns.push_back(VB2_CALL_FN(in.type));
        break;
      }
      default: ok = false;
    }
    o << "      " << V(d) << " = rv; " << N(d) << " = rnull; " << E(d) << " = rerr;\n    }\n";
  }
};

struct Kernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t fn = nullptr;
};

static std::mutex g_mu;
static std::unordered_map<std::string, std::shared_ptr<Kernel>> g_cache;  // value null = known not to compile
static bool g_warned = false;

static std::string cache_key(const vb2_program* p, int n_instrs, const vb2_column* cols, int ncols, bool filter, const vb2_output* outs, int nouts) {
  std::ostringstream k;
  k << (filter ? "F" : "P") << p->filter_reg << ";";
  for (int i = 0; i < n_instrs; ++i) {
    const vb2_instr& in = p->instrs[i];
    k << in.op << "," << in.type << "," << in.dst << "," << in.a << "," << in.b << "," << in.c;
    if (in.op == VB2_OP_CONST) k << "c" << p->consts[in.a].type << p->consts[in.a].is_null;
    if (in.op == VB2_OP_LIKE || in.op == VB2_OP_STRCMP) k << "c" << p->consts[in.b].is_null;
    k << ";";
  }
  k << "|";
  for (int c = 0; c < ncols; ++c) k << cols[c].type << cols[c].encoding << (cols[c].nulls ? 1 : 0) << (cols[c].dict_nulls ? 1 : 0) << ",";
  k << "|";
  for (int i = 0; i < nouts; ++i) k << outs[i].reg << ":" << outs[i].type << ",";
  return k.str();
}

static std::string generate(const vb2_program* p, int n_instrs, const vb2_column* cols, bool filter, const vb2_output* outs, int nouts, bool* ok) {
  Gen g;
  g.prog = p;
  g.cols = cols;
  int nregs = 1;
  for (int i = 0; i < n_instrs; ++i) nregs = std::max(nregs, p->instrs[i].dst + 1);
  nregs = std::max(nregs, static_cast<int>(p->n_regs));
  std::ostringstream& o = g.o;
  o << "extern \"C\" __global__ void __launch_bounds__(256) vb2_jit(const __grid_constant__ JitArgs a) {\n";
  o << "  const long long nwords = (a.n + 31) >> 5;\n  const int lane = threadIdx.x & 31;\n";
  o << "  const long long warp_global = ((long long)blockIdx.x * 256 + threadIdx.x) >> 5;\n  const long long nwarps = ((long long)gridDim.x * 256) >> 5;\n";
  o << "  for (long long w = warp_global; w < nwords; w += nwarps) {\n    const long long k = (w << 5) + lane;\n    const bool live = k < a.n;\n";
  o << "    int errcode = 0;\n";
  for (int r = 0; r < nregs; ++r) o << "    uint64_t v" << r << " = 0; bool n" << r << " = false, e" << r << " = false;\n";
  o << "    if (live) {\n    const long long row = a.sel ? a.sel[k] : k;\n";
  for (int i = 0; i < n_instrs && g.ok; ++i) g.instr(p->instrs[i]);
  o << "    }\n";
  if (filter) {
    const int f = p->filter_reg;
    o << "    bool keep = false;\n    if (live) {\n      if (e" << f << ") atomicCAS(a.error_flag, 0, errcode ? errcode : 1);\n";
    o << "      keep = !e" << f << " && !n" << f << " && v" << f << " != 0;\n    }\n";
    o << "    const unsigned word = __ballot_sync(0xffffffffu, keep);\n    if (lane == 0) a.sel_bits[w] = word;\n";
  } else {
    for (int i = 0; i < nouts; ++i) {
      const int r = outs[i].reg;
      o << "    {\n      bool valid = false;\n      if (live) {\n        if (e" << r << ") atomicCAS(a.error_flag, 0, errcode ? errcode : 1);\n";
      o << "        valid = !e" << r << " && !n" << r << ";\n        const uint64_t v = valid ? v" << r << " : 0;\n";
      const std::string ov = "a.outs[" + std::to_string(i) + "].values";
      if (outs[i].type == VB2_INTEGER) o << "        ((int*)" << ov << ")[k] = (int)v;\n";
      else if (outs[i].type == VB2_BOOLEAN) o << "        ((uint8_t*)" << ov << ")[k] = (uint8_t)v;\n";
      else o << "        ((uint64_t*)" << ov << ")[k] = v;\n";
      o << "      }\n      const unsigned word = __ballot_sync(0xffffffffu, valid);\n      if (lane == 0) ((unsigned*)a.outs[" << i << "].nulls)[w] = word;\n    }\n";
    }
  }
  o << "  }\n}\n";
  *ok = g.ok;
  std::string fns;
  for (int id : g.used_fns) {
    DeviceFn f;
    if (device_fn(id, &f)) fns += "// registered device function " + std::to_string(id) + "\n" + f.source + "\n";
  }
  return std::string(kPrelude) + fns + o.str();
}

static std::shared_ptr<Kernel> compile(const std::string& src) {
  Nvrtc& n = nvrtc();
  if (!n.ok) return nullptr;
  nvrtcProgram prog = nullptr;
  const char* hdr_src[] = {kVmOpsSource};
  const char* hdr_name[] = {"vm_ops.inc"};
  if (n.createProgram(&prog, src.c_str(), "vb2_expr.cu", 1, hdr_src, hdr_name) != NVRTC_SUCCESS) return nullptr;
  const std::string d1 = "-DVB2_SIZEOF_COLUMN=" + std::to_string(sizeof(vb2_column)), d2 = "-DVB2_SIZEOF_CONST=" + std::to_string(sizeof(vb2_const)),
                    d3 = "-DVB2_SIZEOF_OUTPUT=" + std::to_string(sizeof(vb2_output)), d4 = "-DVB2_SIZEOF_ARGS=" + std::to_string(sizeof(JitArgs));
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", d1.c_str(), d2.c_str(), d3.c_str(), d4.c_str()};
  const nvrtcResult rc = n.compileProgram(prog, 7, opts);
  if (rc != NVRTC_SUCCESS) {
    if (!g_warned) {
      g_warned = true;
      size_t ls = 0;
      n.getProgramLogSize(prog, &ls);
      std::string log(ls, '\0');
      if (ls) n.getProgramLog(prog, log.data());
      std::fprintf(stderr, "[velox_b200] expression JIT failed to compile; the interpreter kernels run instead.\n%s\n", log.c_str());
      if (std::getenv("VB2_JIT_DUMP")) std::fprintf(stderr, "%s\n", src.c_str());
    }
    n.destroyProgram(&prog);
    return nullptr;
  }
  size_t sz = 0;
  n.getCUBINSize(prog, &sz);
  std::vector<char> cubin(sz);
  n.getCUBIN(prog, cubin.data());
  n.destroyProgram(&prog);
  auto k = std::make_shared<Kernel>();
  if (cudaLibraryLoadData(&k->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (cudaLibraryGetKernel(&k->fn, k->lib, "vb2_jit") != cudaSuccess) { cudaGetLastError(); return nullptr; }
  return k;
}

static int g_enabled = -1;
static bool enabled() {
  if (g_enabled < 0) {
    const char* e = std::getenv("VB2_EXPR_JIT");
    g_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return g_enabled == 1;
}

static unsigned grid_for(int64_t n) {
  const int64_t b = (n + kThreads - 1) / kThreads;
  const int64_t cap = static_cast<int64_t>(device_sm_count()) * 8;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}

// Returns VB2_ERR_UNSUPPORTED when the caller should run the interpreter instead.
int launch(const vb2_program* p, const vb2_column* cols, int ncols, bool filter, const int32_t* sel, int64_t n, uint32_t* sel_bits,
           const vb2_output* outs, int nouts, int32_t* error_flag, cudaStream_t st) {
  if (!enabled()) return VB2_ERR_UNSUPPORTED;
  if (ncols > kMaxCols || p->n_consts > kMaxConsts || nouts > kMaxOuts || p->n_instrs > 256) return VB2_ERR_UNSUPPORTED;
  const int n_instrs = filter ? p->n_filter_instrs : p->n_instrs;
  const std::string key = cache_key(p, n_instrs, cols, ncols, filter, outs, nouts);
  std::shared_ptr<Kernel> k;
  {
    std::lock_guard<std::mutex> lock(g_mu);
    auto it = g_cache.find(key);
    if (it != g_cache.end()) {
      k = it->second;
    } else {
      bool ok = false;
      const std::string src = generate(p, n_instrs, cols, filter, outs, nouts, &ok);
      if (ok) k = compile(src);
      g_cache.emplace(key, k);
    }
  }
  if (!k) return VB2_ERR_UNSUPPORTED;
  static thread_local JitArgs a;
  for (int i = 0; i < ncols; ++i) a.cols[i] = cols[i];
  for (int i = 0; i < p->n_consts; ++i) a.consts[i] = p->consts[i];
  for (int i = 0; i < nouts; ++i) a.outs[i] = outs[i];
  a.n = n;
  a.sel = sel;
  a.sel_bits = sel_bits;
  a.error_flag = error_flag;
  void* params[] = {&a};
  note_launch();
  const cudaError_t e = cudaLaunchKernel(reinterpret_cast<const void*>(k->fn), dim3(grid_for(n)), dim3(kThreads), params, 0, st);
  if (e != cudaSuccess) return fail_msg(VB2_ERR_CUDA, cudaGetErrorString(e));
  return VB2_OK;
}

}  // namespace jit
}  // namespace vb2

extern "C" {

int32_t vb2k_register_device_function(const char* entry, const char* cuda_source, int32_t ret_type, const int32_t* arg_types, int32_t nargs) {
  using namespace vb2::jit;
  if (!entry || !cuda_source || nargs < 1 || nargs > 3 || !ctype_of(ret_type)) return -VB2_ERR_INVALID;
  DeviceFn f;
  f.entry = entry;
  f.source = cuda_source;
  f.ret = ret_type;
  f.nargs = nargs;
  for (int i = 0; i < nargs; ++i) {
    if (!ctype_of(arg_types[i])) return -VB2_ERR_INVALID;
    f.args[i] = arg_types[i];
  }
  std::lock_guard<std::mutex> lock(g_fn_mu);
  g_fns.push_back(std::move(f));  // ids are never reused: cached kernels of an overwritten function stay valid for old programs
  return static_cast<int32_t>(g_fns.size()) - 1;
}
int32_t vb2k_device_function_count(void) {
  std::lock_guard<std::mutex> lock(vb2::jit::g_fn_mu);
  return static_cast<int32_t>(vb2::jit::g_fns.size());
}

void vb2k_set_expression_jit(int32_t enabled) { vb2::jit::g_enabled = enabled ? 1 : 0; }

// Compiles (without launching) the kernels of a program for the given column layout: 1 = JIT
// kernel available, 0 = the interpreter will run. No GPU needed: used by the CPU test suite.
int32_t vb2k_expression_jit_compiles(const vb2_program* prog, const vb2_column* cols, int32_t ncols, int32_t filter, const vb2_output* outs,
                                     int32_t nouts, char* source_out, int32_t source_len) {
  using namespace vb2::jit;
  bool ok = false;
  const std::string src = generate(prog, filter ? prog->n_filter_instrs : prog->n_instrs, cols, filter != 0, outs, nouts, &ok);
  if (source_out && source_len > 0) std::snprintf(source_out, source_len, "%s", src.c_str());
  if (!ok) return 0;
  Nvrtc& n = nvrtc();
  if (!n.ok) return 0;
  nvrtcProgram p = nullptr;
  const char* hdr_src[] = {kVmOpsSource};
  const char* hdr_name[] = {"vm_ops.inc"};
  if (n.createProgram(&p, src.c_str(), "vb2_expr.cu", 1, hdr_src, hdr_name) != NVRTC_SUCCESS) return 0;
  const std::string d1 = "-DVB2_SIZEOF_COLUMN=" + std::to_string(sizeof(vb2_column)), d2 = "-DVB2_SIZEOF_CONST=" + std::to_string(sizeof(vb2_const)),
                    d3 = "-DVB2_SIZEOF_OUTPUT=" + std::to_string(sizeof(vb2_output)), d4 = "-DVB2_SIZEOF_ARGS=" + std::to_string(sizeof(JitArgs));
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", d1.c_str(), d2.c_str(), d3.c_str(), d4.c_str()};
  const nvrtcResult rc = n.compileProgram(p, 7, opts);
  if (rc != NVRTC_SUCCESS && source_out && source_len > 0) {
    size_t ls = 0;
    n.getProgramLogSize(p, &ls);
    std::string log(ls, '\0');
    if (ls) n.getProgramLog(p, log.data());
    std::snprintf(source_out, source_len, "%s", log.c_str());
  }
  n.destroyProgram(&p);
  return rc == NVRTC_SUCCESS ? 1 : 0;
}

}  // extern "C"
