"""High-cardinality GROUP BY in shared-memory slices (csrc/slice_agg.cu — BASELINE.json configs[4]: the two partition levels
with their histograms, sketch and staged scatters, and the per-slice shared-memory aggregation) compiled FOR THE HOST —
kernels AND the host-side orchestration of vb2k_slice_agg_partition / vb2k_slice_agg_finish, text taken from the .cu file —
and run under the lock-step emulation of tests/host_emulator.py. What the host build replaces: `<<<...>>>` launches become
emulated launches, the handful of CUDA runtime calls become memset / memcpy, and the seven inline-PTX shared-memory
accessors (ld / atom / red on 32-bit shared addresses) become the same operations on a host buffer. Everything else — tile
walks, digit staging, reservations, probing, the two-halves integer sums, chunked output reservation — is the device code.
Checked against a Python group-by: keys, counts, integer sums and MIN / MAX exact, DOUBLE sums within rounding; BIGINT
inputs outside the int32 range take the 64-bit add; an overflowing sum raises error 1. No GPU needed."""
import ctypes as C
import math
import re

import numpy as np
import pytest

from host_emulator import between, build, source

SUM_F64, SUM_I64, COUNT, MIN_I64, MAX_I64 = 1, 2, 3, 6, 7
EMPTY = 0xFFFFFFFFFFFFFFFF

HOST_SHARED_ACCESSORS = r"""
// ---- host stand-ins for the inline-PTX accessors: the "32-bit shared address" is an offset into the table's buffer ----
static uint8_t* g_shared_base = nullptr;
static inline size_t __cvta_generic_to_shared(const void* p) {
  g_shared_base = const_cast<uint8_t*>(static_cast<const uint8_t*>(p)) - 64;  // every thread of the block passes the same pointer
  return 64;
}
template <class T> static inline T* at_shared(uint32_t addr) { return reinterpret_cast<T*>(g_shared_base + addr); }
static inline uint64_t lds_volatile_u64(uint32_t addr) { return __atomic_load_n(at_shared<unsigned long long>(addr), __ATOMIC_SEQ_CST); }
static inline uint64_t atoms_cas_u64(uint32_t addr, uint64_t expect, uint64_t desired) {
  unsigned long long e = expect;
  __atomic_compare_exchange_n(at_shared<unsigned long long>(addr), &e, static_cast<unsigned long long>(desired), false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return e;
}
static inline uint32_t atoms_add_u32(uint32_t addr, uint32_t x) { return __atomic_fetch_add(at_shared<uint32_t>(addr), x, __ATOMIC_SEQ_CST); }
static inline void reds_add_u32(uint32_t addr, uint32_t x) { __atomic_fetch_add(at_shared<uint32_t>(addr), x, __ATOMIC_SEQ_CST); }
static inline uint64_t atoms_add_u64(uint32_t addr, uint64_t x) { return __atomic_fetch_add(at_shared<unsigned long long>(addr), static_cast<unsigned long long>(x), __ATOMIC_SEQ_CST); }
static inline void reds_add_f64(uint32_t addr, double x) { atomicAdd(at_shared<double>(addr), x); }
static inline void reds_minmax_s64(uint32_t addr, int64_t x, bool is_min) {
  if (is_min) atomicMin(at_shared<long long>(addr), static_cast<long long>(x));
  else atomicMax(at_shared<long long>(addr), static_cast<long long>(x));
}
static inline void smem_minmax_f64(uint32_t addr, double v, bool is_min) {
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
"""

BODY_HEAD = r"""
// ---- CUDA runtime and helper stand-ins of the host build ----
typedef void* cudaStream_t;
enum { cudaMemcpyDeviceToHost = 2, cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
static inline int cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return 0; }
static inline int cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { std::memcpy(d, s, n); return 0; }
static inline int cudaStreamSynchronize(cudaStream_t) { return 0; }
static inline int cudaGetLastError() { return 0; }
template <class K> static inline int cudaFuncSetAttribute(K, int, int) { return 0; }
#define VB2_CUDA_OK(x) do { (void)(x); } while (0)
static inline int fail_msg(int code, const char*) { return code; }
static inline int device_sm_count() { return 2; }
static inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
struct ulonglong2 { unsigned long long x, y; };
static inline ulonglong2 make_ulonglong2(unsigned long long x, unsigned long long y) { return ulonglong2{x, y}; }
template <class T> static inline T counted(T g) { return g; }
"""


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    common, text = source("common.cuh"), source("slice_agg.cu")
    kernels = between(text, "constexpr int kPT = 512;", "}  // namespace\n}  // namespace vb2")
    entry = between(text, "int32_t vb2k_slice_agg_hll_registers(void)", '}  // extern "C"')
    # the inline-PTX accessors -> host stand-ins; the opaque register copy of the shared base -> nothing; dynamic shared memory -> static arrays
    ptx_begin = kernels.index("// ---- shared-memory access by 32-bit shared address")
    ptx_end = kernels.index("// One accumulator update in shared memory")
    kernels = kernels[:ptx_begin] + HOST_SHARED_ACCESSORS + kernels[ptx_end:]
    kernels = re.sub(r'\n\s*asm volatile\("mov\.u32 %0, %0;" : "\+r"\(skey_s\)\);[^\n]*', "\n", kernels)
    assert "asm volatile" not in kernels
    kernels = kernels.replace("extern __shared__ __align__(16) uint8_t smem[];", "static __attribute__((aligned(16))) uint8_t smem[224 * 1024];")
    launch = re.compile(r"(\b\w+(?:<[^;<>]*>)?)<<<counted\((.*?)\), ([^,]+), ([^,]+), st>>>\((.*?)\);")
    both = kernels + "\n// ---- the C entry points (host orchestration) ----\n" + entry
    both, n = launch.subn(r"launch(\2, \3, [&] { \1(\5); });", both)
    assert n >= 10 and "<<<" not in both, n
    body = (BODY_HEAD + "// ---- common.cuh: hash mixers, warp reductions ----\n" +
            between(common, "__host__ __device__ __forceinline__ uint64_t twang_mix64", "__device__ __forceinline__ uint64_t hash_f64") +
            between(common, "__device__ __forceinline__ double warp_sum(double v)", "}  // namespace vb2") +
            "// ---- slice_agg.cu ----\n" + both + "\n}  // namespace vb2_on_host\n"
            'extern "C" {\n'
            "size_t h_workspace(int64_t n, int ncols) { return vb2_on_host::vb2k_slice_agg_workspace(n, ncols); }\n"
            "int64_t h_output_rows(int64_t d) { return vb2_on_host::vb2k_slice_agg_output_rows(d); }\n"
            "int h_partition(const vb2_slice_chunk* c, int nc, int ncols, int64_t n, void* ws, size_t wsb, int32_t* hll) {\n"
            "  return vb2_on_host::vb2k_slice_agg_partition(c, nc, ncols, n, ws, wsb, hll, nullptr); }\n"
            "int h_finish(int64_t n, int ncols, int64_t distinct, const vb2_slice_op* ops, int nops, int rw, const uint64_t* init, uint64_t* rows, int64_t cap,\n"
            "             int64_t* groups, int64_t* reserved, int32_t* err, int32_t* ovf, void* ws, size_t wsb) {\n"
            "  return vb2_on_host::vb2k_slice_agg_finish(n, ncols, distinct, ops, nops, rw, init, rows, cap, groups, reserved, err, ovf, ws, wsb, nullptr); }\n"
            "}\n")
    L = build(tmp_path_factory.mktemp("slice_on_host"), "slice", body)
    L.h_workspace.restype = C.c_size_t
    L.h_output_rows.restype = C.c_int64
    return L


class Chunk(C.Structure):
    _fields_ = [("norm_keys", C.c_void_p), ("raw_keys", C.c_void_p), ("key_min", C.c_int64), ("cols", C.c_void_p * 3), ("rows", C.c_int64)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("col", C.c_int32), ("word", C.c_int32)]


def run_slices(host, chunks, ncols, ops, row_words, row_init, distinct_estimate):
    """chunks: list of (raw keys int64, key_min, [payload arrays]). Returns (group rows ndarray [groups, row_words], error, overflowed slices)."""
    total = sum(len(c[0]) for c in chunks)
    carr = (Chunk * len(chunks))()
    for i, (keys, kmin, cols) in enumerate(chunks):
        carr[i].raw_keys, carr[i].key_min, carr[i].rows = keys.ctypes.data, kmin, len(keys)
        for c, a in enumerate(cols):
            carr[i].cols[c] = a.ctypes.data
    wsb = int(host.h_workspace(C.c_int64(total), ncols))
    ws = np.zeros(wsb + 64, dtype=np.uint8)
    hll = np.zeros(4096, dtype=np.int32)
    assert host.h_partition(carr, len(chunks), ncols, C.c_int64(total), ws.ctypes.data_as(C.c_void_p), C.c_size_t(wsb), hll.ctypes.data_as(C.c_void_p)) == 0
    cap = 1
    while cap < max(16, int(host.h_output_rows(C.c_int64(min(total, distinct_estimate))))):
        cap <<= 1
    rows = np.zeros(cap * row_words, dtype=np.uint64)
    for r in range(cap):
        rows[r * row_words:(r + 1) * row_words] = row_init
    rows[0::row_words] = np.uint64(EMPTY)
    words = np.zeros(4, dtype=np.int64)
    flags = np.zeros(4, dtype=np.int32)
    oarr = (Op * len(ops))(*[Op(*o) for o in ops])
    init = np.array(row_init, dtype=np.uint64)
    rc = host.h_finish(C.c_int64(total), ncols, C.c_int64(distinct_estimate), oarr, len(ops), row_words, init.ctypes.data_as(C.c_void_p),
                       rows.ctypes.data_as(C.c_void_p), C.c_int64(cap), words.ctypes.data_as(C.c_void_p), words[1:].ctypes.data_as(C.c_void_p),
                       flags.ctypes.data_as(C.c_void_p), flags[1:].ctypes.data_as(C.c_void_p), ws.ctypes.data_as(C.c_void_p), C.c_size_t(wsb))
    assert rc == 0
    table = rows.reshape(cap, row_words)
    occupied = table[table[:, 0] != np.uint64(EMPTY)]
    assert int(words[0]) == len(occupied) and int(words[1]) >= len(occupied)  # groups written; rows reserved in chunks
    return occupied, int(flags[0]), int(flags[1])


i64 = lambda w: int(np.array([w], dtype=np.uint64).view(np.int64)[0])       # noqa: E731
f64 = lambda w: float(np.array([w], dtype=np.uint64).view(np.float64)[0])   # noqa: E731


@pytest.mark.parametrize("n,distinct,estimate", [(9_000, 1500, 1500), (3_000, 150, 300_000)])
def test_slice_pipeline_matches_a_python_group_by(host, n, distinct, estimate):
    """Two input chunks, one BIGINT and one DOUBLE payload. An estimate of 300 K groups sizes level 2 (256 x 2 slices, most of
    them empty over 150 real keys: every emulated slice costs four 512-thread barriers, so the case is kept small); the
    other case stays at one level."""
    rng = np.random.default_rng(estimate % 97)
    keys = rng.integers(0, distinct, n) * 7919 - 3_000_000
    kmin = int(keys.min())
    v0 = rng.integers(-1000, 1000, n).astype(np.int64)
    v1 = np.round(rng.normal(0, 10, n), 2)
    half = n // 2 + 17
    chunks = [(np.ascontiguousarray(keys[:half]), kmin, [np.ascontiguousarray(v0[:half]), np.ascontiguousarray(v1[:half])]),
              (np.ascontiguousarray(keys[half:]), kmin, [np.ascontiguousarray(v0[half:]), np.ascontiguousarray(v1[half:])])]
    row_words = 8  # [key | sum v0 | count | max v0 | sum v1 | min v0 | untouched x2]
    init = [0, 0, 0, np.uint64(np.iinfo(np.int64).min + 2**64), 0, np.uint64(np.iinfo(np.int64).max), 12345, 0]
    ops = [(SUM_I64, 0, 1), (COUNT, -1, 2), (MAX_I64, 0, 3), (SUM_F64, 1, 4), (MIN_I64, 0, 5)]
    got, err, ovf = run_slices(host, chunks, 2, ops, row_words, init, estimate)
    assert err == 0 and ovf == 0
    want = {}
    for k, a, b in zip(keys, v0, v1):
        g = want.setdefault(int(k) - kmin + 1, [0, 0, None, 0.0, None])
        g[0] += int(a)
        g[1] += 1
        g[2] = int(a) if g[2] is None else max(g[2], int(a))
        g[3] += float(b)
        g[4] = int(a) if g[4] is None else min(g[4], int(a))
    assert len(got) == len(want)
    for r in got:
        g = want[int(r[0])]
        assert i64(r[1]) == g[0] and int(r[2]) == g[1] and i64(r[3]) == g[2] and i64(r[5]) == g[4]
        assert math.isclose(f64(r[4]), g[3], rel_tol=1e-12, abs_tol=1e-9)
        assert int(r[6]) == 12345 and int(r[7]) == 0  # untouched words come from row_init


def test_wide_integer_sums_and_overflow_on_the_slice_path(host):
    """The shape of tests/test_slice_agg_gpu.py::test_slice_path_wide_integer_sums_and_overflow: values outside the int32 range
    take the 64-bit shared-memory add beside the two-halves path of small ones; a sum leaving int64 raises error 1."""
    rng = np.random.default_rng(9)
    n, distinct = 6_000, 1200
    keys = rng.integers(0, distinct, n) * 31 - 5_000_000
    mags = rng.choice([1, 1000, 2**31 - 1, 2**31, 2**40, 2**45], n)
    v = ((rng.integers(0, 1000, n) + 1) * mags * rng.choice([-1, 1], n)).astype(np.int64)
    kmin = int(keys.min())
    got, err, ovf = run_slices(host, [(np.ascontiguousarray(keys), kmin, [v])], 1, [(SUM_I64, 0, 1), (COUNT, -1, 2)], 4, [0, 0, 0, 0], distinct)
    assert err == 0 and ovf == 0
    want = {}
    for k, a in zip(keys, v):
        g = want.setdefault(int(k) - kmin + 1, [0, 0])
        g[0] += int(a)
        g[1] += 1
    assert {int(r[0]): [i64(r[1]), int(r[2])] for r in got} == want
    big = np.full(n, 2**62, dtype=np.int64)
    big[: n // 2] = 1
    _, err, _ = run_slices(host, [(np.ascontiguousarray(keys), kmin, [big])], 1, [(SUM_I64, 0, 1), (COUNT, -1, 2)], 4, [0, 0, 0, 0], distinct)
    assert err == 1
