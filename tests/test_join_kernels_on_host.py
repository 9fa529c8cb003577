"""The hash-join kernels (csrc/hash_join.cu: build with chains, count / emit and the one-pass unique-key probe, array and
hash table modes) and the keyed key-id kernel of the keyed (kHash) join mode (csrc/hash_agg.cu) compiled FOR THE HOST and
run thread by thread on the CPU — any sequential order of the threads is one valid interleaving of their atomics; the
trailing warp ballots are resolved in two passes — against numpy: the same (probe row, build row) pairs, in probe order,
that HashTable::listJoinResults enumerates (exec/HashTable.cpp:2133-2350), NULL keys never matching
(exec/HashBuild.cpp:475-479). No GPU needed; the GPU suite runs the kernels on the device through the operators."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SHIM = r"""
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "velox_b200_kernels.h"
namespace vb2_on_host {
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static Dim3 threadIdx, blockIdx, gridDim, blockDim;
static int phase = 0;
static size_t ballot_at = 0;
static std::vector<unsigned> masks;
static inline unsigned __ballot_sync(unsigned, bool p) {
  const size_t i = ballot_at++;
  if (phase == 0) {
    if (masks.size() <= i) masks.resize(i + 1, 0u);
    if (p) masks[i] |= 1u << (threadIdx.x & 31u);
    return 0u;
  }
  return masks[i];
}
static inline int atomicCAS(int* p, int e, int v) { const int o = *p; if (o == e) *p = v; return o; }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long e, unsigned long long v) { const auto o = *p; if (o == e) *p = v; return o; }
static inline int atomicExch(int* p, int v) { const int o = *p; *p = v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const auto o = *p; *p += v; return o; }
static inline void __threadfence() {}
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline int64_t __mul64hi(int64_t a, int64_t b) { return static_cast<int64_t>((static_cast<__int128>(a) * b) >> 64); }
static inline int64_t warp_sum(int64_t v) { return v; }  // only feeds an optional group counter (not requested here)
using std::isnan;
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(...)
#define __restrict__
#include "vm_ops.inc"
constexpr uint64_t kNullHash = 1;
// ---- common.cuh: hash mixers ----
%(mixers)s
// ---- hash_join.cu: find_slot and the build / probe kernels ----
%(join)s
// ---- hash_agg.cu: column decoding and the keyed table ----
%(decode)s
%(keyed)s
}  // namespace vb2_on_host

using namespace vb2_on_host;
template <class F>
static void launch(int64_t n, bool ballots, F&& kernel) {  // blocks of 256 threads, enough to give every row a thread
  blockDim.x = 256;
  gridDim.x = static_cast<unsigned>((n + 255) / 256 > 0 ? (n + 255) / 256 : 1);
  for (unsigned b = 0; b < gridDim.x; ++b) {
    blockIdx.x = b;
    for (unsigned warp = 0; warp < 8; ++warp) {
      masks.clear();
      for (phase = ballots ? 0 : 1; phase < 2; ++phase)
        for (unsigned lane = 0; lane < 32; ++lane) {
          threadIdx.x = warp * 32 + lane;
          ballot_at = 0;
          kernel();
        }
    }
  }
}
extern "C" {
void h_join_build(const vb2_join_table* t, const uint64_t* keys, const uint64_t* valid, int64_t n, int32_t* flags) {
  launch(n, false, [&] { join_build_kernel(*t, keys, valid, n, flags); });
}
void h_join_count(const vb2_join_table* t, const uint64_t* keys, const uint64_t* valid, int64_t n, int32_t* counts) {
  launch(n, false, [&] { join_probe_count_kernel(*t, keys, valid, n, counts); });
}
void h_join_emit(const vb2_join_table* t, const uint64_t* keys, const uint64_t* valid, int64_t n, const int64_t* offsets, int32_t* pr, int32_t* br) {
  launch(n, false, [&] { join_probe_emit_kernel(*t, keys, valid, n, offsets, pr, br); });
}
void h_join_unique(const vb2_join_table* t, const uint64_t* keys, const uint64_t* valid, int64_t n, uint32_t* hit_bits, int32_t* hits) {
  launch(n, true, [&] { join_probe_unique_kernel(*t, keys, valid, n, hit_bits, hits); });
}
void h_keyed_ids(const vb2_group_table* t, const vb2_column* cols, int nkeys, int64_t n, int insert, uint64_t* ids, uint32_t* valid, int32_t* err) {
  KeyedCols kc{};
  kc.n = nkeys;
  for (int k = 0; k < nkeys; ++k) kc.c[k] = cols[k];
  launch(n, true, [&] { keyed_key_ids_kernel(*t, kc, n, insert, ids, valid, nullptr, err); });
}
}
"""


class JoinTable(C.Structure):
    _fields_ = [("mode", C.c_int32), ("pad", C.c_int32), ("key_min", C.c_int64), ("capacity", C.c_int64), ("keys", C.c_void_p), ("head", C.c_void_p),
                ("next", C.c_void_p), ("build_rows", C.c_int64)]


class GroupTable(C.Structure):
    _fields_ = [("rows", C.c_void_p), ("capacity", C.c_int64), ("row_words", C.c_int32), ("hash_mode", C.c_int32)]


def _between(text, begin, end):
    b = text.index(begin)
    return text[b:text.index(end, b)]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    csrc = os.path.join(ROOT, "velox_b200", "csrc")
    common = open(os.path.join(csrc, "common.cuh")).read()
    join = open(os.path.join(csrc, "hash_join.cu")).read()
    agg = open(os.path.join(csrc, "hash_agg.cu")).read()
    parts = {
        "mixers": _between(common, "__host__ __device__ __forceinline__ uint64_t twang_mix64", "__device__ __forceinline__ uint64_t hash_f64"),
        "join": _between(join, "__device__ __forceinline__ int64_t find_slot", "// ---- exclusive scan of int32 counts"),
        "decode": _between(agg, "__device__ __forceinline__ bool decode_row2", "__global__ void normalize_keys_kernel"),
        "keyed": _between(agg, "struct KeyedCols {", "__global__ void group_update_keyed_kernel") +
                 _between(agg, "// Slot of a key tuple already in a keyed table", "// Re-inserts groups into a (bigger) keyed table"),
    }
    d = tmp_path_factory.mktemp("join_on_host")
    src = d / "join.cpp"
    src.write_text(SHIM % parts)
    so = d / "libjoin.so"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-w", "-I", os.path.join(ROOT, "include"), "-I", csrc,
                           "-o", str(so), str(src)])
    return C.CDLL(str(so))


def _bits(valid):
    n = len(valid)
    words = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    for i, v in enumerate(valid):
        if v:
            words[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    return words


def _reference_pairs(build_keys, build_valid, probe_keys, probe_valid):
    """(probe row, build row) pairs in probe order; per probe row the build rows of its key (any order: compared as sets)."""
    rows = {}
    for r, (k, ok) in enumerate(zip(build_keys, build_valid)):
        if ok:
            rows.setdefault(int(k), []).append(r)
    return [(p, sorted(rows.get(int(k), []))) for p, (k, ok) in enumerate(zip(probe_keys, probe_valid)) if ok and int(k) in rows]


def _run_join(host, mode, build_keys, build_valid, probe_keys, probe_valid):
    n, m = len(probe_keys), len(build_keys)
    t = JoinTable()
    t.mode = mode
    bk = np.ascontiguousarray(build_keys, dtype=np.uint64)
    pk = np.ascontiguousarray(probe_keys, dtype=np.uint64)
    if mode == 0:
        t.key_min = 0
        t.capacity = int(max(int(bk.max()) if m else 0, 1)) + 1
        keys = None
    else:
        cap = 16
        while cap < 2 * m + 16:
            cap <<= 1
        t.capacity = cap
        keys = np.full(cap, 0xFFFFFFFFFFFFFFFF, dtype=np.uint64)
        t.keys = keys.ctypes.data
    head = np.zeros(t.capacity, dtype=np.int32)
    nxt = np.zeros(m + 1, dtype=np.int32)
    t.head, t.next, t.build_rows = head.ctypes.data, nxt.ctypes.data, m
    bv, pv = _bits(build_valid), _bits(probe_valid)
    flags = np.zeros(4, dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)
    host.h_join_build(C.byref(t), P(bk), P(bv), C.c_int64(m), P(flags))
    assert flags[0] == 0
    counts = np.zeros(n + 1, dtype=np.int32)
    host.h_join_count(C.byref(t), P(pk), P(pv), C.c_int64(n), P(counts))
    offsets = np.zeros(n + 1, dtype=np.int64)
    offsets[1:] = np.cumsum(counts[:n])
    total = int(offsets[n])
    pr, br = np.full(total + 1, -1, dtype=np.int32), np.full(total + 1, -1, dtype=np.int32)
    host.h_join_emit(C.byref(t), P(pk), P(pv), C.c_int64(n), P(offsets), P(pr), P(br))
    hit_bits = np.zeros((n + 31) // 32 + 2, dtype=np.uint32)
    hits = np.full(n + 1, -7, dtype=np.int32)
    host.h_join_unique(C.byref(t), P(pk), P(pv), C.c_int64(n), P(hit_bits), P(hits))
    return bool(flags[1]), counts[:n], pr[:total], br[:total], hit_bits, hits[:n]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("dup", [False, True])
def test_join_build_and_probe_kernels(host, mode, dup):
    rng = np.random.default_rng(mode * 2 + dup)
    n, m = 1500, 260
    space = 400 if mode == 0 else 2**40            # array mode: a dense key range; hash mode: sparse 40-bit keys
    if dup:
        bk = rng.integers(0, space, m) if mode == 0 else rng.choice(rng.integers(0, space, 90), m)
    else:
        bk = rng.permutation(space)[:m] if mode == 0 else rng.permutation(5000)[:m] * 2**28
    pk = np.concatenate([rng.choice(bk, n // 2), rng.integers(0, space, n - n // 2)])
    rng.shuffle(pk)
    bvalid, pvalid = rng.random(m) > 0.06, rng.random(n) > 0.06
    saw_dups, counts, pr, br, hit_bits, hits = _run_join(host, mode, bk, bvalid, pk, pvalid)
    want = _reference_pairs(bk, bvalid, pk, pvalid)
    valid_build_keys = [int(k) for k, ok in zip(bk, bvalid) if ok]
    assert saw_dups == (len(set(valid_build_keys)) != len(valid_build_keys))
    # count / emit: pairs in probe order, every chain complete
    assert [int(c) for c in counts] == [len(dict(want).get(p, [])) for p in range(n)]
    got, at = [], 0
    for p, c in enumerate(counts):
        if c:
            assert np.all(pr[at:at + c] == p)
            got.append((p, sorted(int(x) for x in br[at:at + c])))
            at += c
    assert got == want
    # one-pass probe: the ballot bitmap marks the matching probe rows, hits hold one matching build row (the only one without duplicates)
    matched = {p for p, _ in want}
    assert {i for i in range(n) if (int(hit_bits[i >> 5]) >> (i & 31)) & 1} == matched
    for p, rows in want:
        assert int(hits[p]) in rows
    assert all(int(hits[p]) == -1 for p in range(n) if p not in matched)


def test_keyed_key_ids_kernel(host):
    """Key tuples -> ids through a keyed table: equal tuples share an id, tuples with a NULL column get none, the probe side
    finds exactly the build side's tuples; DOUBLE columns compare by canonical bits (NaN = NaN, -0 = +0)."""
    from velox_b200.vector import BIGINT, DOUBLE, flat_vector
    rng = np.random.default_rng(5)
    m, n = 400, 1200
    nan = float("nan")

    def columns(count, seed):
        r = np.random.default_rng(seed)
        a = [None if r.random() < 0.05 else int(v) * 2**40 for v in r.integers(0, 9, count)]
        d = [None if r.random() < 0.05 else float(v) for v in r.choice([0.0, -0.0, nan, 1.5, -2.25, 7.0], count)]
        return a, d, [flat_vector(BIGINT, a), flat_vector(DOUBLE, d)]

    ba, bd, bcols = columns(m, 1)
    pa, pd_, pcols = columns(n, 2)
    cap = 1024
    row_words = 4
    rows = np.zeros(cap * row_words, dtype=np.uint64)
    rows[0::row_words] = np.uint64(0xFFFFFFFFFFFFFFFF)
    t = GroupTable(rows.ctypes.data, cap, row_words, 2)
    P = lambda a: a.ctypes.data_as(C.c_void_p)

    def ids_of(cols, count, insert):
        from velox_b200.kernels import CColumn
        arr = (CColumn * 2)(*[c.to_c() for c in cols])
        ids = np.zeros(count + 1, dtype=np.uint64)
        valid = np.zeros((count + 31) // 32 + 2, dtype=np.uint32)
        err = np.zeros(2, dtype=np.int32)
        host.h_keyed_ids(C.byref(t), arr, 2, C.c_int64(count), insert, P(ids), P(valid), P(err))
        assert err[0] == 0
        return [int(ids[i]) if (int(valid[i >> 5]) >> (i & 31)) & 1 else None for i in range(count)]

    def canon(a, d):
        if a is None or d is None:
            return None
        return (a, "nan" if d != d else (0.0 if d == 0.0 else d))

    bids = ids_of(bcols, m, 1)
    by_tuple = {}
    for i in range(m):
        key = canon(ba[i], bd[i])
        assert (bids[i] is None) == (key is None)
        if key is not None:
            assert by_tuple.setdefault(key, bids[i]) == bids[i]
    assert len(set(by_tuple.values())) == len(by_tuple)  # different tuples, different ids
    pids = ids_of(pcols, n, 0)
    for i in range(n):
        key = canon(pa[i], pd_[i])
        assert pids[i] == (by_tuple.get(key) if key is not None else None), (i, key)
