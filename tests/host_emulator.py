"""TEST INFRASTRUCTURE — a small lock-step emulation of the CUDA execution model for running extracted kernel source on
the CPU: every CUDA thread of a block is an OS thread, blocks run one after another, __syncthreads and the warp
collectives (__shfl_up_sync / __shfl_xor_sync / __match_any_sync / __ballot_sync / __syncwarp) are barriers over per-warp
exchange slots, atomics are real atomics, `__shared__` variables are function-local statics (one block at a time).
Used by the *_on_host tests, which splice kernel text taken verbatim from velox_b200/csrc/*.cu after PRELUDE."""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "velox_b200", "csrc")

PRELUDE = r"""
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>
#include "velox_b200_kernels.h"
namespace vb2_on_host {
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static thread_local Dim3 threadIdx;
static Dim3 blockIdx, gridDim, blockDim;   // one block runs at a time
static std::unique_ptr<std::barrier<>> block_barrier;
static std::vector<std::unique_ptr<std::barrier<>>> warp_barrier;
static long long exchange[32][32];          // [warp][lane]
static inline void __syncthreads() { block_barrier->arrive_and_wait(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { warp_barrier[threadIdx.x >> 5]->arrive_and_wait(); }
template <class T, class Pick>
static inline T warp_exchange(T v, Pick pick) {  // every lane publishes v, then reads the lane pick(own lane) names
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  exchange[w][l] = bits;
  warp_barrier[w]->arrive_and_wait();
  const long long got = exchange[w][pick(l) & 31u];
  warp_barrier[w]->arrive_and_wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int delta) { return warp_exchange(v, [delta](unsigned l) { return l >= static_cast<unsigned>(delta) ? l - delta : l; }); }
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) { return warp_exchange(v, [mask](unsigned l) { return l ^ static_cast<unsigned>(mask); }); }
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) { return warp_exchange(v, [src](unsigned) { return static_cast<unsigned>(src); }); }
static inline unsigned __match_any_sync(unsigned, unsigned v) {
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  exchange[w][l] = v;
  warp_barrier[w]->arrive_and_wait();
  unsigned m = 0;
  for (unsigned o = 0; o < 32; ++o)
    if (static_cast<unsigned>(exchange[w][o]) == v) m |= 1u << o;
  warp_barrier[w]->arrive_and_wait();
  return m;
}
static inline unsigned __ballot_sync(unsigned, bool p) {
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  exchange[w][l] = p ? 1 : 0;
  warp_barrier[w]->arrive_and_wait();
  unsigned m = 0;
  for (unsigned o = 0; o < 32; ++o)
    if (exchange[w][o]) m |= 1u << o;
  warp_barrier[w]->arrive_and_wait();
  return m;
}
static inline bool __any_sync(unsigned m, bool p) { return __ballot_sync(m, p) != 0; }
static inline bool __all_sync(unsigned m, bool p) { return __ballot_sync(m, p) == 0xffffffffu; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs(static_cast<int>(v)); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
static inline int atomicCAS(int* p, int e, int v) { __atomic_compare_exchange_n(p, &e, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return e; }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long e, unsigned long long v) {
  __atomic_compare_exchange_n(p, &e, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return e;
}
static inline double atomicAdd(double* p, double v) {
  unsigned long long* q = reinterpret_cast<unsigned long long*>(p);
  unsigned long long old = __atomic_load_n(q, __ATOMIC_SEQ_CST);
  for (;;) {
    double d;
    std::memcpy(&d, &old, 8);
    const double sum = d + v;
    unsigned long long want;
    std::memcpy(&want, &sum, 8);
    if (__atomic_compare_exchange_n(q, &old, want, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) return d;
  }
}
static inline long long atomicMin(long long* p, long long v) {
  long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v < old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline long long atomicMax(long long* p, long long v) {
  long long old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline int __clzll(long long v) { return v ? __builtin_clzll(static_cast<unsigned long long>(v)) : 64; }
static inline void __threadfence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline void __threadfence_system() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
using std::min;
using std::max;
static inline int64_t __mul64hi(int64_t a, int64_t b) { return static_cast<int64_t>((static_cast<__int128>(a) * b) >> 64); }
using std::isnan;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__
#define __shared__ static
#define __align__(x)
constexpr int kWarp = 32;
#include "vm_ops.inc"
template <class F>
static void launch(unsigned grid, unsigned threads, F&& kernel) {
  gridDim.x = grid;
  blockDim.x = threads;
  for (unsigned b = 0; b < grid; ++b) {
    blockIdx.x = b;
    block_barrier = std::make_unique<std::barrier<>>(threads);
    warp_barrier.clear();
    for (unsigned w = 0; w < (threads + 31) / 32; ++w) warp_barrier.push_back(std::make_unique<std::barrier<>>(32));
    std::vector<std::thread> ts;
    for (unsigned t = 0; t < threads; ++t)
      ts.emplace_back([&, t] {
        threadIdx.x = t;
        kernel();
      });
    for (auto& th : ts) th.join();
  }
}
"""


def between(text, begin, end):
    b = text.index(begin)
    return text[b:text.index(end, b)]


def source(name):
    with open(os.path.join(CSRC, name)) as f:
        return f.read()


def build(tmpdir, name, body):
    """Compiles PRELUDE + body (kernel text inside namespace vb2_on_host, then the extern "C" drivers) into a shared library."""
    src = tmpdir / f"{name}.cpp"
    src.write_text(PRELUDE + body)
    so = tmpdir / f"lib{name}.so"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-Wl,-Bsymbolic", "-ffp-contract=off", "-w", "-I", os.path.join(ROOT, "include"),
                           "-I", CSRC, "-o", str(so), str(src)])
    return C.CDLL(str(so))
