// NVRTC behind dlopen, shared by the expression JIT (expr_jit.cu) and the pipeline JIT
// (fused_jit.cu). The library must load on hosts without CUDA, so nothing links against libnvrtc.
#pragma once
#include <dlfcn.h>
#include <nvrtc.h>

#include <mutex>
#include <string>
#include <vector>

namespace vb2 {
namespace jit {

struct Nvrtc {
  void* handle = nullptr;
  nvrtcResult (*createProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  nvrtcResult (*compileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  nvrtcResult (*getCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*getCUBIN)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*getProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  nvrtcResult (*getProgramLog)(nvrtcProgram, char*) = nullptr;
  nvrtcResult (*destroyProgram)(nvrtcProgram*) = nullptr;
  nvrtcResult (*addNameExpression)(nvrtcProgram, const char*) = nullptr;
  nvrtcResult (*getLoweredName)(nvrtcProgram, const char*, const char**) = nullptr;
  bool ok = false;
};

inline Nvrtc& nvrtc() {
  static Nvrtc n;
  static std::once_flag once;
  std::call_once(once, [] {
    const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "libnvrtc.so"};
    for (const char* nm : names) {
      n.handle = dlopen(nm, RTLD_NOW | RTLD_LOCAL);
      if (n.handle) break;
    }
    if (!n.handle) return;
    auto sym = [&](const char* s) { return dlsym(n.handle, s); };
    n.createProgram = reinterpret_cast<decltype(n.createProgram)>(sym("nvrtcCreateProgram"));
    n.compileProgram = reinterpret_cast<decltype(n.compileProgram)>(sym("nvrtcCompileProgram"));
    n.getCUBINSize = reinterpret_cast<decltype(n.getCUBINSize)>(sym("nvrtcGetCUBINSize"));
    n.getCUBIN = reinterpret_cast<decltype(n.getCUBIN)>(sym("nvrtcGetCUBIN"));
    n.getProgramLogSize = reinterpret_cast<decltype(n.getProgramLogSize)>(sym("nvrtcGetProgramLogSize"));
    n.getProgramLog = reinterpret_cast<decltype(n.getProgramLog)>(sym("nvrtcGetProgramLog"));
    n.destroyProgram = reinterpret_cast<decltype(n.destroyProgram)>(sym("nvrtcDestroyProgram"));
    n.addNameExpression = reinterpret_cast<decltype(n.addNameExpression)>(sym("nvrtcAddNameExpression"));
    n.getLoweredName = reinterpret_cast<decltype(n.getLoweredName)>(sym("nvrtcGetLoweredName"));
    n.ok = n.createProgram && n.compileProgram && n.getCUBINSize && n.getCUBIN && n.getProgramLogSize && n.getProgramLog && n.destroyProgram &&
           n.addNameExpression && n.getLoweredName;
  });
  return n;
}

}  // namespace jit
}  // namespace vb2
