"""The stable partition order of B200PartitionedOutput (csrc/hash_partition.cu: per-block histograms, the parallel offsets
scan, the warp-ranked stable scatter — HashPartitionFunction::partition followed by the per-destination grouping of
exec/PartitionedOutput.cpp) compiled FOR THE HOST and run under a small lock-step emulation: every CUDA thread of a
block is an OS thread, __syncthreads and the warp collectives (__shfl_up_sync, __match_any_sync) are barriers over
exchange slots. Result against numpy's stable argsort, partition ids given or computed on the fly from a BIGINT key
(folly::hasher = twang_mix64, then % partitions — bit-exact with the oracle). No GPU needed."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIM = r"""
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
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
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int delta) {
  const unsigned w = threadIdx.x >> 5, l = threadIdx.x & 31;
  long long bits = 0;
  std::memcpy(&bits, &v, sizeof(T));
  exchange[w][l] = bits;
  warp_barrier[w]->arrive_and_wait();
  const long long got = exchange[w][l >= static_cast<unsigned>(delta) ? l - delta : l];
  warp_barrier[w]->arrive_and_wait();
  T out;
  std::memcpy(&out, &got, sizeof(T));
  return out;
}
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
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
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
constexpr int kWarp = 32;
#include "vm_ops.inc"
// ---- common.cuh: hash mixers ----
%(mixers)s
// ---- hash_partition.cu: the stable partition order ----
%(part)s
}  // namespace vb2_on_host

using namespace vb2_on_host;
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
        vb2_on_host::threadIdx.x = t;
        kernel();
      });
    for (auto& th : ts) th.join();
  }
}
extern "C" void h_partition_order(const uint32_t* ids, const void* key, int is64, int64_t rows, int parts, int64_t* counts, int32_t* order) {
  const PartSrc src{ids, key, is64};
  const int64_t nblocks = (rows + kPartRowsPerBlock - 1) / kPartRowsPerBlock;
  std::vector<int32_t> hist(nblocks * parts);
  std::vector<int64_t> base(nblocks * parts);
  launch(static_cast<unsigned>(nblocks), kPartThreads, [&] { part_hist_kernel(src, rows, parts, hist.data()); });
  launch(1, kOffsetThreads, [&] { part_offsets_kernel(hist.data(), nblocks, parts, counts, base.data()); });
  launch(static_cast<unsigned>(nblocks), kPartThreads, [&] { part_scatter_kernel(src, rows, parts, base.data(), order); });
}
"""


def _between(text, begin, end):
    b = text.index(begin)
    return text[b:text.index(end, b)]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    csrc = os.path.join(ROOT, "velox_b200", "csrc")
    common = open(os.path.join(csrc, "common.cuh")).read()
    part = open(os.path.join(csrc, "hash_partition.cu")).read()
    parts = {
        "mixers": _between(common, "__host__ __device__ __forceinline__ uint64_t twang_mix64", "__device__ __forceinline__ uint64_t hash_f64"),
        "part": _between(part, "constexpr int kPartThreads", "// --- fixed-capacity segments"),
    }
    d = tmp_path_factory.mktemp("partition_on_host")
    src = d / "part.cpp"
    src.write_text(SHIM % parts)
    so = d / "libpart.so"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-pthread", "-Wl,-Bsymbolic", "-w", "-I", os.path.join(ROOT, "include"), "-I", csrc,
                           "-o", str(so), str(src)])
    return C.CDLL(str(so))


P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


@pytest.mark.parametrize("parts", [2, 7, 64])
def test_stable_partition_order_with_given_ids(host, parts):
    rng = np.random.default_rng(parts)
    n = 3 * 4096 + 1234  # four blocks, the last one partial
    ids = rng.integers(0, parts, n).astype(np.uint32)
    ids[100:3000] = 1 % parts  # a long run of one partition
    counts = np.zeros(parts, dtype=np.int64)
    order = np.full(n, -1, dtype=np.int32)
    host.h_partition_order(P(ids), None, 0, C.c_int64(n), parts, P(counts), P(order))
    assert np.array_equal(counts, np.bincount(ids, minlength=parts))
    assert np.array_equal(order, np.argsort(ids, kind="stable").astype(np.int32))


def test_partition_ids_from_a_key_column_match_the_oracle(host):
    """part_id computed on the fly = HashPartitionFunction over one BIGINT key: twang_mix64(key) % partitions
    (exec/HashPartitionFunction.cpp:75-118, exec/VectorHasher.cpp:62-126), the routing a CPU worker would compute."""
    rng = np.random.default_rng(1)
    n, parts = 2 * 4096 + 77, 8
    keys = rng.integers(-2**62, 2**62, n)
    counts = np.zeros(parts, dtype=np.int64)
    order = np.full(n, -1, dtype=np.int32)
    host.h_partition_order(None, P(keys), 1, C.c_int64(n), parts, P(counts), P(order))
    L = pyoracle.lib()
    want_ids = np.array([L.orc_twang_mix64(int(k) & 0xFFFFFFFFFFFFFFFF) % parts for k in keys], dtype=np.int64)
    assert np.array_equal(counts, np.bincount(want_ids, minlength=parts))
    assert np.array_equal(order, np.argsort(want_ids, kind="stable").astype(np.int32))
