"""The general GROUP BY kernels (csrc/hash_agg.cu: key normalisation into value ids — VectorHasher's range mode,
exec/VectorHasher.h:523-585 — and group_update_kernel: find-or-insert of the row's group + every accumulator update,
the role of HashTable::groupProbe + Aggregate::addRawInput, exec/HashTable.cpp:470-523 and
functions/lib/aggregates/SimpleNumericAggregate.h:94-150) compiled FOR THE HOST and run thread by thread on the CPU
against a pure-Python group-by: keys with NULLs (a NULL key is a group, GroupingSet.cpp:448-455), SUM / COUNT / MIN / MAX
over BIGINT and DOUBLE inputs with NULLs, an aggregate mask, a dictionary-wrapped input, hash and array table modes.
No GPU needed."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest

from velox_b200.kernels import CColumn
from velox_b200.vector import BIGINT, DOUBLE, INTEGER, flat_vector

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EMPTY = 0xFFFFFFFFFFFFFFFF
SUM_F64, SUM_I64, COUNT, MIN_F64, MAX_F64, MIN_I64, MAX_I64 = 1, 2, 3, 4, 5, 6, 7

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
static inline void __threadfence() {}
static inline int atomicCAS(int* p, int e, int v) { const int o = *p; if (o == e) *p = v; return o; }
static inline unsigned long long atomicCAS(unsigned long long* p, unsigned long long e, unsigned long long v) { const auto o = *p; if (o == e) *p = v; return o; }
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { const auto o = *p; *p += v; return o; }
static inline double atomicAdd(double* p, double v) { const double o = *p; *p = o + v; return o; }
static inline long long atomicMin(long long* p, long long v) { const auto o = *p; if (v < o) *p = v; return o; }
static inline long long atomicMax(long long* p, long long v) { const auto o = *p; if (v > o) *p = v; return o; }
static inline int __clzll(long long v) { return v ? __builtin_clzll(static_cast<unsigned long long>(v)) : 64; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline int64_t __mul64hi(int64_t a, int64_t b) { return static_cast<int64_t>((static_cast<__int128>(a) * b) >> 64); }
static inline int64_t warp_sum(int64_t v) { return v; }  // no cross-lane sum in this harness: the new-group counter is not requested
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
// ---- hash_agg.cu: key normalisation ----
%(norm)s
// ---- hash_agg.cu: accumulator updates, group table, group_update_kernel ----
%(update)s
// ---- hash_agg.cu: keyed (kHash) tables ----
%(keyed)s
}  // namespace vb2_on_host

using namespace vb2_on_host;
template <class F>
static void launch(int64_t n, bool ballots, F&& kernel) {
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
void h_normalize(const vb2_column* cols, int ncols, const int64_t* mins, const uint64_t* mults, const uint64_t* ranges, int nulls_invalid, int check_ranges,
                 int64_t n, uint64_t* out, uint32_t* valid) {
  NormArgs a{};
  a.n = ncols;
  for (int k = 0; k < ncols; ++k) { a.c[k] = cols[k]; a.mins[k] = mins[k]; a.mults[k] = mults[k]; a.ranges[k] = ranges ? ranges[k] : 0; }
  a.nulls_invalid = nulls_invalid;
  a.check_ranges = check_ranges;
  launch(n, valid != nullptr, [&] { normalize_keys_kernel(a, nullptr, n, out, valid); });
}
void h_group_update(const vb2_group_table* t, const uint64_t* keys, const uint64_t* valid, int64_t n, const vb2_agg_update* aggs, int naggs,
                    int64_t* num_groups, int32_t* error_flag) {
  AggArgs args{};
  args.n = naggs;
  for (int i = 0; i < naggs; ++i) args.a[i] = aggs[i];
  launch(n, false, [&] { group_update_kernel(*t, keys, valid, n, args, num_groups, error_flag); });
}
void h_group_update_keyed(const vb2_group_table* t, const vb2_column* cols, int nkeys, int64_t n, const vb2_agg_update* aggs, int naggs, int32_t* error_flag) {
  KeyedCols kc{};
  kc.n = nkeys;
  for (int k = 0; k < nkeys; ++k) kc.c[k] = cols[k];
  AggArgs args{};
  args.n = naggs;
  for (int i = 0; i < naggs; ++i) args.a[i] = aggs[i];
  launch(n, false, [&] { group_update_keyed_kernel(*t, kc, n, args, nullptr, error_flag); });
}
}
"""


class GroupTable(C.Structure):
    _fields_ = [("rows", C.c_void_p), ("capacity", C.c_int64), ("row_words", C.c_int32), ("hash_mode", C.c_int32)]


class AggUpdate(C.Structure):
    _fields_ = [("kind", C.c_int32), ("input_type", C.c_int32), ("input", C.c_void_p), ("nulls", C.c_void_p), ("mask", C.c_void_p), ("indices", C.c_void_p),
                ("base_nulls", C.c_void_p), ("acc_word", C.c_int32), ("nonnull_word", C.c_int32)]


def _between(text, begin, end):
    b = text.index(begin)
    return text[b:text.index(end, b)]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    csrc = os.path.join(ROOT, "velox_b200", "csrc")
    common = open(os.path.join(csrc, "common.cuh")).read()
    agg = open(os.path.join(csrc, "hash_agg.cu")).read()
    parts = {
        "mixers": _between(common, "__host__ __device__ __forceinline__ uint64_t twang_mix64", "__device__ __forceinline__ uint64_t hash_f64"),
        "norm": _between(agg, "constexpr int kMaxNormCols", "__global__ void minmax_kernel"),
        "update": _between(agg, "__device__ __forceinline__ double input_as_f64", "// Array-mode tables from a handful to a few thousand groups"),
        "keyed": _between(agg, "struct KeyedCols {", "// Slot of a key tuple already in a keyed table"),
    }
    d = tmp_path_factory.mktemp("agg_on_host")
    src = d / "agg.cpp"
    src.write_text(SHIM % parts)
    so = d / "libagg.so"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-ffp-contract=off", "-w", "-I", os.path.join(ROOT, "include"),
                           "-I", csrc, "-o", str(so), str(src)])
    return C.CDLL(str(so))


def _validity(values, as_bool=False):
    """u64 bitmap: bit i set when values[i] is not None (as_bool: when it is true)."""
    n = len(values)
    words = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    for i, v in enumerate(values):
        if (bool(v) if as_bool else v is not None):
            words[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    return words


P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


@pytest.mark.parametrize("hash_mode", [1, 0])
def test_normalize_and_group_update(host, hash_mode):
    rng = np.random.default_rng(3 + hash_mode)
    n = 3000
    k0 = [None if rng.random() < 0.07 else int(v) for v in rng.integers(-4, 5, n)]       # BIGINT key, range 9 (+ NULL id)
    k1 = [None if rng.random() < 0.07 else int(v) for v in rng.integers(100, 106, n)]    # INTEGER key, range 6 (+ NULL id)
    x = [None if rng.random() < 0.1 else float(v) for v in np.round(rng.normal(0, 50, n), 3)]
    y = [None if rng.random() < 0.1 else int(v) for v in rng.integers(-1000, 1000, n)]
    mask = rng.random(n) < 0.7
    base = np.round(rng.normal(0, 5, 40), 2)                                            # dictionary-wrapped DOUBLE input
    idx = rng.integers(0, 40, n).astype(np.int32)
    kc = [flat_vector(BIGINT, k0), flat_vector(INTEGER, k1)]
    cols = (CColumn * 2)(*[c.to_c() for c in kc])
    mins = np.array([-4, 100], dtype=np.int64)
    ranges = np.array([9 + 1, 6 + 1], dtype=np.uint64)  # ids 1..range, 0 = NULL
    mults = np.array([7, 1], dtype=np.uint64)
    keys = np.zeros(n + 1, dtype=np.uint64)
    host.h_normalize(cols, 2, P(mins), P(mults), None, 0, 0, C.c_int64(n), P(keys), None)
    for i in range(n):
        id0 = 0 if k0[i] is None else k0[i] + 4 + 1
        id1 = 0 if k1[i] is None else k1[i] - 100 + 1
        assert int(keys[i]) == id0 * 7 + id1
    # the join form: a NULL key column clears the row's valid bit, out-of-range ids too
    valid = np.zeros((n + 31) // 32 + 2, dtype=np.uint32)
    tight = np.array([5, 7], dtype=np.uint64)  # ids of column 0 above 4 are out of range now
    host.h_normalize(cols, 2, P(mins), P(mults), P(tight), 1, 1, C.c_int64(n), P(keys), P(valid))
    for i in range(n):
        ok = k0[i] is not None and k1[i] is not None and (k0[i] + 4 + 1) < 5
        assert ((int(valid[i >> 5]) >> (i & 31)) & 1) == int(ok), i
    host.h_normalize(cols, 2, P(mins), P(mults), None, 0, 0, C.c_int64(n), P(keys), None)  # group-by keys again

    # group table: word 0 key / occupancy, accumulators [sum x, nn x, sum y, nn y, count(*), min y, max x, masked sum y, nn, sum dict, nn] ...
    row_words = 12
    capacity = 256 if hash_mode else 70 * 1 + 8  # array mode: the packed key space (10 * 7 = 70 ids)
    rows = np.zeros(capacity * row_words, dtype=np.uint64)
    init = np.zeros(row_words, dtype=np.uint64)
    init[0] = EMPTY if hash_mode else 0
    init[6] = np.uint64(np.iinfo(np.int64).max)           # min identity
    init[7] = np.array([-math.inf]).view(np.uint64)[0]     # max identity
    for r in range(capacity):
        rows[r * row_words:(r + 1) * row_words] = init
    t = GroupTable(rows.ctypes.data, capacity, row_words, hash_mode)
    xa = np.array([0.0 if v is None else v for v in x])
    ya = np.array([0 if v is None else v for v in y], dtype=np.int64)
    xv, yv, mv = _validity(x), _validity(y), _validity(mask, as_bool=True)
    aggs = (AggUpdate * 7)(
        AggUpdate(SUM_F64, DOUBLE, xa.ctypes.data, xv.ctypes.data, None, None, None, 1, 2),
        AggUpdate(SUM_I64, BIGINT, ya.ctypes.data, yv.ctypes.data, None, None, None, 3, 4),
        AggUpdate(COUNT, BIGINT, None, None, None, None, None, 5, -1),
        AggUpdate(MIN_I64, BIGINT, ya.ctypes.data, yv.ctypes.data, None, None, None, 6, -1),
        AggUpdate(MAX_F64, DOUBLE, xa.ctypes.data, xv.ctypes.data, None, None, None, 7, -1),
        AggUpdate(SUM_I64, BIGINT, ya.ctypes.data, yv.ctypes.data, mv.ctypes.data, None, None, 8, 9),
        AggUpdate(SUM_F64, DOUBLE, base.ctypes.data, None, None, idx.ctypes.data, None, 10, 11))
    err = np.zeros(2, dtype=np.int32)
    host.h_group_update(C.byref(t), P(keys), None, C.c_int64(n), aggs, 7, None, P(err))
    assert err[0] == 0
    # reference: plain python over the same normalized keys
    want = {}
    for i in range(n):
        g = want.setdefault(int(keys[i]), {"sx": 0.0, "nx": 0, "sy": 0, "ny": 0, "c": 0, "miny": None, "maxx": None, "msy": 0, "mn": 0, "sd": 0.0, "nd": 0})
        g["c"] += 1
        if x[i] is not None:
            g["sx"] += x[i]
            g["nx"] += 1
            g["maxx"] = x[i] if g["maxx"] is None else max(g["maxx"], x[i])
        if y[i] is not None:
            g["sy"] += y[i]
            g["ny"] += 1
            g["miny"] = y[i] if g["miny"] is None else min(g["miny"], y[i])
            if mask[i]:
                g["msy"] += y[i]
                g["mn"] += 1
        g["sd"] += float(base[idx[i]])
        g["nd"] += 1
    table = rows.reshape(capacity, row_words)
    if hash_mode:
        got_rows = {int(r[0]): r for r in table if int(r[0]) != EMPTY}
    else:
        got_rows = {slot: r for slot, r in enumerate(table) if int(r[0]) != 0}
    assert set(got_rows) == set(want)
    f64 = lambda w: float(np.array([w], dtype=np.uint64).view(np.float64)[0])  # noqa: E731
    i64 = lambda w: int(np.array([w], dtype=np.uint64).view(np.int64)[0])      # noqa: E731
    for k, g in want.items():
        r = got_rows[k]
        assert math.isclose(f64(r[1]), g["sx"], rel_tol=1e-12, abs_tol=1e-9) and int(r[2]) == g["nx"]
        assert i64(r[3]) == g["sy"] and int(r[4]) == g["ny"] and int(r[5]) == g["c"]
        assert (i64(r[6]) == g["miny"]) if g["miny"] is not None else (i64(r[6]) == np.iinfo(np.int64).max)
        assert (f64(r[7]) == g["maxx"]) if g["maxx"] is not None else (f64(r[7]) == -math.inf)
        assert i64(r[8]) == g["msy"] and int(r[9]) == g["mn"]
        assert math.isclose(f64(r[10]), g["sd"], rel_tol=1e-12, abs_tol=1e-9) and int(r[11]) == g["nd"]


def test_sum_overflow_sets_the_error_flag(host):
    """SUM(BIGINT) is checked (functions/prestosql/aggregates/SumAggregate.cpp): an overflowing group raises error 1."""
    n = 64
    keys = np.ones(n + 1, dtype=np.uint64)
    vals = np.full(n, 2**62, dtype=np.int64)
    rows = np.zeros(8 * 4, dtype=np.uint64)
    rows[0::4] = np.uint64(EMPTY)
    t = GroupTable(rows.ctypes.data, 8, 4, 1)
    aggs = (AggUpdate * 1)(AggUpdate(SUM_I64, BIGINT, vals.ctypes.data, None, None, None, None, 1, 2))
    err = np.zeros(2, dtype=np.int32)
    host.h_group_update(C.byref(t), P(keys), None, C.c_int64(n), aggs, 1, None, P(err))
    assert err[0] == 1


def test_keyed_group_update(host):
    """The keyed (kHash) group table (exec/HashTable.cpp:1751-1838): rows store their key columns; DOUBLE keys group by canonical
    bits (every NaN one group, -0 with +0), a wide BIGINT key beside it, NULL keys form groups of their own
    (GroupingSet.cpp:448-455)."""
    rng = np.random.default_rng(8)
    n = 2500
    nan = float("nan")
    a = [None if rng.random() < 0.06 else int(v) * 2**41 - 2**55 for v in rng.integers(0, 6, n)]
    d = [None if rng.random() < 0.06 else float(v) for v in rng.choice([0.0, -0.0, nan, 1.5, -2.25], n)]
    y = rng.integers(-100, 100, n).astype(np.int64)
    key_columns = [flat_vector(BIGINT, a), flat_vector(DOUBLE, d)]  # kept: to_c() points into buffers these objects own
    cols = (CColumn * 2)(*[c.to_c() for c in key_columns])
    nkeys, row_words, capacity = 2, 8, 128          # [state | key words x2 | NULL mask | sum | count | pad x2]
    rows = np.zeros(capacity * row_words, dtype=np.uint64)
    rows[0::row_words] = np.uint64(EMPTY)
    t = GroupTable(rows.ctypes.data, capacity, row_words, 2)
    aggs = (AggUpdate * 2)(AggUpdate(SUM_I64, BIGINT, y.ctypes.data, None, None, None, None, 4, -1), AggUpdate(COUNT, BIGINT, None, None, None, None, None, 5, -1))
    err = np.zeros(2, dtype=np.int32)
    host.h_group_update_keyed(C.byref(t), cols, nkeys, C.c_int64(n), aggs, 2, P(err))
    assert err[0] == 0
    want = {}
    for i in range(n):
        dk = None if d[i] is None else ("nan" if d[i] != d[i] else (0.0 if d[i] == 0.0 else d[i]))
        g = want.setdefault((a[i], dk), [0, 0])
        g[0] += int(y[i])
        g[1] += 1
    table = rows.reshape(capacity, row_words)
    got = {}
    for r in table:
        if int(r[0]) == EMPTY:
            continue
        assert int(r[0]) & 1  # published
        nullmask = int(r[3])
        ka = None if nullmask & 1 else int(np.array([r[1]], dtype=np.uint64).view(np.int64)[0])
        kd = None
        if not nullmask & 2:
            v = float(np.array([r[2]], dtype=np.uint64).view(np.float64)[0])
            kd = "nan" if v != v else v
        got[(ka, kd)] = [int(np.array([r[4]], dtype=np.uint64).view(np.int64)[0]), int(r[5])]
    assert got == want
