// Pipeline JIT: the expression templates of fused_scan.cuh are instantiated at run time, with NVRTC,
// for plan shapes that have no ahead-of-time specialisation. The canonical signature the planner
// prints for a Filter/Project/Aggregation chain (host/expr_compiler.cpp fusedSignature) is parsed
// back into the template type it describes —
//     F:and(between(i0,pi0,pi1),lt(f1,pf0));P:multiply(f2,f1)
//  -> Pipeline<And<Between<ColI<0>,PI<0>,PI<1>>, Lt<ColF<1>,PF<0>>>, TypeList<Multiply<ColF<2>,ColF<1>>>, -1>
// — and the SAME kernels the registered pipelines use (TMA-staged scan, direct-load scan, filter
// bitmap, gather-aggregate) are compiled for it, lazily per variant, and launched by the same
// launch logic (fused_scan.cu). So any supported expression shape gets the whole scan -> filter ->
// project -> aggregate chain in one HBM-bound kernel, not only TPC-H's.
// (The reference's GPU prototype generates one-thread-per-row kernels from its plans,
// velox/experimental/wave/exec/WaveGen.cpp:820; it, too, compiles with NVRTC, wave/jit.)
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "fused_scan.cuh"
#include "jit_common.h"
#include "common_str.h"      // kCommonCuhSource
#include "fused_scan_str.h"  // kFusedScanCuhSource
#include "vm_ops_str.h"      // kVmOpsSource

static_assert(VB2_FUSED_MAX_COLS == 8 && VB2_FUSED_MAX_PARAMS == 12 && VB2_FUSED_MAX_KEYS == 2,
              "common.cuh spells these constants out for NVRTC: keep both in step");

namespace vb2 {
namespace fx {

namespace {

// ---- signature -> template type -----------------------------------------------------------------
struct SigParser {
  const std::string& s;
  size_t p = 0;
  bool ok = true;
  uint32_t fmask = 0, imask = 0, lmask = 0;
  bool uses_join = false;

  static bool ident_char(char c) { return (c >= 'a' && c <= 'z') || (c >= '0' && c <= '9') || c == '_'; }

  // one expression starting at p; appends the C++ type text
  void expr(std::string& out) {
    size_t b = p;
    while (p < s.size() && ident_char(s[p])) ++p;
    const std::string id = s.substr(b, p - b);
    if (id.empty()) { ok = false; return; }
    if (p < s.size() && s[p] == '(') {
      static const std::map<std::string, std::string> fn = {
          {"plus", "Plus"}, {"minus", "Minus"}, {"multiply", "Multiply"}, {"divide", "Divide"}, {"lt", "Lt"}, {"lte", "Lte"}, {"gt", "Gt"},
          {"gte", "Gte"}, {"eq", "Eq"}, {"neq", "Neq"}, {"between", "Between"}, {"and", "And"}, {"switch", "Switch"}};
      auto it = fn.find(id);
      if (it == fn.end()) { ok = false; return; }
      out += it->second + "<";
      ++p;
      int nargs = 0;
      for (;;) {
        if (nargs++) out += ",";
        expr(out);
        if (!ok) return;
        if (p < s.size() && s[p] == ',') { ++p; continue; }
        if (p < s.size() && s[p] == ')') { ++p; break; }
        ok = false;
        return;
      }
      out += ">";
      return;
    }
    if (id == "true") { out += "True"; return; }
    if (id == "joinflag") { out += "JoinFlag"; uses_join = true; return; }
    auto number = [&](size_t from) -> int {
      if (from >= id.size()) { ok = false; return 0; }
      int v = 0;
      for (size_t i = from; i < id.size(); ++i) {
        if (id[i] < '0' || id[i] > '9') { ok = false; return 0; }
        v = v * 10 + (id[i] - '0');
      }
      return v;
    };
    if (id[0] == 'p' && id.size() >= 3 && (id[1] == 'f' || id[1] == 'i' || id[1] == 'l')) {
      const int k = number(2);
      if (k >= VB2_FUSED_MAX_PARAMS) ok = false;
      out += std::string(id[1] == 'f' ? "PF<" : (id[1] == 'i' ? "PI<" : "PL<")) + std::to_string(k) + ">";
      return;
    }
    if (id[0] == 'f' || id[0] == 'i' || id[0] == 'l') {
      const int c = number(1);
      if (c >= VB2_FUSED_MAX_COLS) { ok = false; return; }
      (id[0] == 'f' ? fmask : (id[0] == 'i' ? imask : lmask)) |= 1u << c;
      out += std::string(id[0] == 'f' ? "ColF<" : (id[0] == 'i' ? "ColI<" : "ColL<")) + std::to_string(c) + ">";
      return;
    }
    ok = false;
  }
};

struct Parsed {
  std::string filter_type, pipeline_type;  // C++ type text
  PipelineDesc desc;
  bool filter_only = false;
  bool ok = false;
};

Parsed parse_signature(const std::string& sig) {
  Parsed out;
  if (sig.compare(0, 2, "F:") != 0) return out;
  const size_t pp = sig.find(";P:");
  if (pp == std::string::npos) return out;
  if (sig.find(";C:") != std::string::npos) return out;  // compaction pipelines are ahead-of-time only
  const std::string f = sig.substr(2, pp - 2);
  std::string rest = sig.substr(pp + 3);
  int join_col = -1;
  const size_t jp = rest.find(";J:l");
  if (jp != std::string::npos) {
    join_col = std::atoi(rest.c_str() + jp + 4);
    rest = rest.substr(0, jp);
  }
  SigParser fp{f};
  std::string ftype;
  fp.expr(ftype);
  if (!fp.ok || fp.p != f.size()) return out;
  out.filter_type = ftype;
  out.desc.has_filter = ftype != "True";
  out.desc.ffmask = fp.fmask;
  out.desc.fimask = fp.imask;
  out.desc.flmask = fp.lmask;
  out.desc.fmask = fp.fmask;
  out.desc.imask = fp.imask;
  out.desc.lmask = fp.lmask;
  if (rest.empty()) {
    if (!out.desc.has_filter || join_col >= 0) return out;
    out.filter_only = true;
    out.ok = true;
    return out;
  }
  std::string projs;
  int np = 0;
  size_t b = 0;
  int depth = 0;
  for (size_t i = 0; i <= rest.size(); ++i) {
    if (i < rest.size() && rest[i] == '(') ++depth;
    if (i < rest.size() && rest[i] == ')') --depth;
    if (i == rest.size() || (rest[i] == '|' && depth == 0)) {
      const std::string one = rest.substr(b, i - b);
      SigParser pr{one};
      std::string t;
      pr.expr(t);
      if (!pr.ok || pr.p != one.size()) return out;
      out.desc.fmask |= pr.fmask;
      out.desc.imask |= pr.imask;
      out.desc.lmask |= pr.lmask;
      if (pr.uses_join && join_col < 0) return out;
      projs += (np ? "," : "") + t;
      ++np;
      b = i + 1;
    }
  }
  if (np == 0 || np > VB2_FUSED_MAX_COLS) return out;
  if (join_col >= VB2_FUSED_MAX_COLS) return out;
  if (join_col >= 0) out.desc.lmask |= 1u << join_col;
  out.desc.nproj = np;
  out.desc.join = join_col >= 0;
  out.pipeline_type = "Pipeline<" + ftype + ",TypeList<" + projs + ">," + std::to_string(join_col) + ">";
  out.ok = true;
  return out;
}

// ---- compile -----------------------------------------------------------------------------------------
struct JitKernel {
  cudaLibrary_t lib = nullptr;
  cudaKernel_t fn = nullptr;
};

std::string name_expression(const Parsed& p, KernelKind kind, int maxg, bool key64) {
  const std::string key = key64 ? "long long" : "int";
  const std::string P = "vb2::fx::JitPipeline";
  switch (kind) {
    case KernelKind::kTma: return "&vb2::fx::fused_scan_agg_tma_kernel<" + P + "," + std::to_string(maxg) + "," + key + ">";
    case KernelKind::kDirect: return "&vb2::fx::fused_scan_agg_kernel<" + P + "," + std::to_string(maxg) + ",2,false," + key + ">";
    case KernelKind::kFilterBits:
      return "&vb2::fx::fused_filter_bits_tma_kernel<vb2::fx::JitFilterView," + std::to_string(filter_tile_rows_for(p.desc.ffmask, p.desc.fimask, p.desc.flmask)) + ">";
    case KernelKind::kGather: return "&vb2::fx::fused_gather_agg_kernel<" + P + "," + std::to_string(maxg) + "," + key + ">";
  }
  return "";
}

std::string source_of(const Parsed& p) {
  std::string src = "#include \"fused_scan.cuh\"\nnamespace vb2 { namespace fx {\n";
  src += "struct JitFilterView { static constexpr uint32_t fmask = " + p.filter_type + "::fmask, imask = " + p.filter_type + "::imask, lmask = " +
         p.filter_type + "::lmask; using F = " + p.filter_type + "; };\n";
  if (!p.filter_only) src += "using JitPipeline = " + p.pipeline_type + ";\n";
  src += "} }\n";
  return src;
}

bool g_warned = false;

// Compiles one kernel variant. cubin_out only: no GPU needed (CPU test); otherwise loads it.
bool compile_variant(const Parsed& p, const std::string& name_expr, std::vector<char>* cubin_out, std::string* lowered, std::string* log_out) {
  jit::Nvrtc& n = jit::nvrtc();
  if (!n.ok) { if (log_out) *log_out = "NVRTC is not available"; return false; }
  const std::string src = source_of(p);
  nvrtcProgram prog = nullptr;
  const char* hdr_src[] = {kFusedScanCuhSource, kCommonCuhSource, kVmOpsSource};
  const char* hdr_name[] = {"fused_scan.cuh", "common.cuh", "vm_ops.inc"};
  if (n.createProgram(&prog, src.c_str(), "vb2_fused.cu", 3, hdr_src, hdr_name) != NVRTC_SUCCESS) return false;
  n.addNameExpression(prog, name_expr.c_str());
  const char* opts[] = {"--gpu-architecture=sm_100a", "--std=c++17", "--fmad=false", "-lineinfo"};
  const nvrtcResult rc = n.compileProgram(prog, 4, opts);
  if (rc != NVRTC_SUCCESS) {
    size_t ls = 0;
    n.getProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) n.getProgramLog(prog, log.data());
    if (log_out) *log_out = log;
    if (!g_warned && std::getenv("VB2_JIT_DUMP")) {
      g_warned = true;
      std::fprintf(stderr, "[velox_b200] pipeline JIT failed to compile %s\n%s\n%s\n", name_expr.c_str(), log.c_str(), src.c_str());
    }
    n.destroyProgram(&prog);
    return false;
  }
  const char* low = nullptr;
  if (n.getLoweredName(prog, name_expr.c_str(), &low) != NVRTC_SUCCESS || !low) { n.destroyProgram(&prog); return false; }
  if (lowered) *lowered = low;
  size_t sz = 0;
  n.getCUBINSize(prog, &sz);
  cubin_out->resize(sz);
  n.getCUBIN(prog, cubin_out->data());
  n.destroyProgram(&prog);
  return true;
}

struct JitPipelineState {
  Parsed parsed;
  std::mutex mu;
  std::map<int, std::shared_ptr<JitKernel>> variants;  // null = known not to compile
  const void* get(KernelKind kind, int maxg, bool key64) {
    if (kind == KernelKind::kFilterBits) { maxg = 0; key64 = false; }
    if (parsed.filter_only && kind != KernelKind::kFilterBits) return nullptr;
    if ((kind == KernelKind::kFilterBits || kind == KernelKind::kGather) && !parsed.desc.has_filter) return nullptr;
    const int key = static_cast<int>(kind) * 1000 + maxg * 10 + (key64 ? 1 : 0);
    std::lock_guard<std::mutex> lock(mu);
    auto it = variants.find(key);
    if (it == variants.end()) {
      std::shared_ptr<JitKernel> k;
      std::vector<char> cubin;
      std::string lowered;
      if (compile_variant(parsed, name_expression(parsed, kind, maxg, key64), &cubin, &lowered, nullptr)) {
        k = std::make_shared<JitKernel>();
        if (cudaLibraryLoadData(&k->lib, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess ||
            cudaLibraryGetKernel(&k->fn, k->lib, lowered.c_str()) != cudaSuccess) {
          cudaGetLastError();
          k = nullptr;
        }
      }
      it = variants.emplace(key, k).first;
    }
    return it->second ? reinterpret_cast<const void*>(it->second->fn) : nullptr;
  }
};

const void* jit_kernels(void* self, KernelKind kind, int maxg, bool key64) { return static_cast<JitPipelineState*>(self)->get(kind, maxg, key64); }

int g_enabled = -1;
bool enabled() {
  if (g_enabled < 0) {
    const char* e = std::getenv("VB2_PIPELINE_JIT");
    g_enabled = (e && e[0] == '0') ? 0 : 1;
  }
  return g_enabled == 1;
}

}  // namespace

// Registers a run-time pipeline for `signature` (kernels compile at first use); -1 when the
// signature is outside what the templates express.
int jit_pipeline(const std::string& signature) {
  if (!enabled() || !jit::nvrtc().ok) return -1;
  Parsed p = parse_signature(signature);
  if (!p.ok) return -1;
  auto* state = new JitPipelineState();  // lives as long as the registry (process lifetime)
  state->parsed = p;
  Entry e;
  e.signature = signature;
  e.nproj = p.desc.nproj;
  e.join = p.desc.join;
  e.desc = p.desc;
  e.kernels = &jit_kernels;
  e.self = state;
  return register_pipeline(e);
}

}  // namespace fx
}  // namespace vb2

extern "C" {

void vb2k_set_pipeline_jit(int32_t enabled) { vb2::fx::g_enabled = enabled ? 1 : 0; }

// Compiles (does not load or launch; no GPU needed) one kernel of the pipeline a signature describes:
// kind 0 = TMA scan-aggregate, 1 = direct-load scan-aggregate, 2 = filter bitmap, 3 = gather-aggregate.
// Returns 1 on success; log_out receives the generated type or the compiler log.
int32_t vb2k_pipeline_jit_compiles(const char* signature, int32_t kind, int32_t max_groups, int32_t key64, char* log_out, int32_t log_len) {
  using namespace vb2::fx;
  Parsed p = parse_signature(signature ? signature : "");
  if (!p.ok) {
    if (log_out && log_len > 0) std::snprintf(log_out, log_len, "signature outside the template grammar");
    return 0;
  }
  std::vector<char> cubin;
  std::string lowered, log;
  const bool ok = compile_variant(p, name_expression(p, static_cast<KernelKind>(kind), max_groups, key64 != 0), &cubin, &lowered, &log);
  if (log_out && log_len > 0) std::snprintf(log_out, log_len, "%s", ok ? (p.filter_only ? p.filter_type.c_str() : p.pipeline_type.c_str()) : log.c_str());
  return ok ? 1 : 0;
}

}  // extern "C"
