"""The block-private shared-memory group table for low-cardinality GROUP BY (csrc/hash_agg.cu group_update_smem_kernel: every
block accumulates into its own copy of a small array-mode table and merges it into the global table once) compiled FOR THE
HOST and run under the lock-step emulation of tests/host_emulator.py — real concurrent threads, real atomics — against a
plain Python group-by and against group_update_kernel (the global-table path) over the same rows: counts, integer sums and
MIN / MAX exact, DOUBLE sums within rounding (the order of atomic adds is not fixed). No GPU needed."""
import ctypes as C
import math

import numpy as np
import pytest

from host_emulator import between, build, source

BIGINT, DOUBLE = 4, 6
SUM_F64, SUM_I64, COUNT, MIN_I64, MAX_F64 = 1, 2, 3, 6, 5

BODY = r"""
constexpr uint64_t kNullHash = 1;
// ---- common.cuh: hash mixers, warp reductions ----
%(mixers)s
%(reductions)s
// ---- hash_agg.cu: key normalisation ----
%(norm)s
// ---- hash_agg.cu: accumulator updates, the group table, group_update_kernel, group_update_smem_kernel ----
%(update)s
// ---- hash_agg.cu: the serial (input-order) kernel of very small batches and the register-accumulator kernel of tiny tables ----
%(tiny)s
// ---- hash_agg.cu: occupancy and the extraction of result columns ----
%(occupied)s
%(extract)s
}  // namespace vb2_on_host
using namespace vb2_on_host;
extern "C" {
void h_group_update(int smem, const vb2_group_table* t, const uint64_t* keys, int64_t n, const vb2_agg_update* aggs, int naggs, int32_t* error_flag, int blocks) {
  AggArgs args{};
  args.n = naggs;
  for (int i = 0; i < naggs; ++i) args.a[i] = aggs[i];
  if (smem) launch(blocks, 256, [&] { group_update_smem_kernel(*t, keys, nullptr, n, args, error_flag); });
  else launch(blocks, 256, [&] { group_update_kernel(*t, keys, nullptr, n, args, nullptr, error_flag); });
}
// the dispatch of vb2k_group_update for a table of at most 8 groups: an occupancy pass, then one pass per SUM / COUNT
void h_group_update_tiny(const vb2_group_table* t, const uint64_t* keys, int64_t n, const vb2_agg_update* aggs, int naggs, int32_t* error_flag, int blocks) {
  vb2_agg_update none{};
  launch(blocks, 256, [&] { group_update_tiny_kernel<0>(*t, keys, nullptr, n, none, error_flag); });
  for (int i = 0; i < naggs; ++i) {
    const vb2_agg_update u = aggs[i];
    switch (u.kind) {
      case VB2_AGG_SUM_F64: launch(blocks, 256, [&] { group_update_tiny_kernel<VB2_AGG_SUM_F64>(*t, keys, nullptr, n, u, error_flag); }); break;
      case VB2_AGG_SUM_I64: launch(blocks, 256, [&] { group_update_tiny_kernel<VB2_AGG_SUM_I64>(*t, keys, nullptr, n, u, error_flag); }); break;
      default: launch(blocks, 256, [&] { group_update_tiny_kernel<VB2_AGG_COUNT>(*t, keys, nullptr, n, u, error_flag); });
    }
  }
}
// GROUP BY end to end: normalized keys -> group_update_kernel -> group_extract_kernel (one block scans the table for its groups)
void h_group_by(const vb2_column* key_cols, int nkeys, const int64_t* mins, const uint64_t* mults, int64_t n, const vb2_group_table* t,
                const vb2_agg_update* aggs, int naggs, const vb2_extract_col* out_cols, int nout, int32_t* scratch, int64_t* header, int32_t* error_flag) {
  NormArgs na{};
  na.n = nkeys;
  for (int k = 0; k < nkeys; ++k) { na.c[k] = key_cols[k]; na.mins[k] = mins[k]; na.mults[k] = mults[k]; }
  std::vector<uint64_t> keys(n + 1);
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  launch(blocks, 256, [&] { normalize_keys_kernel(na, nullptr, n, keys.data(), nullptr); });
  AggArgs args{};
  args.n = naggs;
  for (int i = 0; i < naggs; ++i) args.a[i] = aggs[i];
  launch(blocks, 256, [&] { group_update_kernel(*t, keys.data(), nullptr, n, args, nullptr, error_flag); });
  ExtractArgs ea{};
  ea.n = nout;
  for (int i = 0; i < nout; ++i) ea.c[i] = out_cols[i];
  launch(1, 1024, [&] { group_extract_kernel(*t, nullptr, 0, scratch, ea, header, error_flag); });
}
void h_group_update_serial(const vb2_group_table* t, const uint64_t* keys, int64_t n, const vb2_agg_update* aggs, int naggs, int32_t* error_flag) {
  AggArgs args{};
  args.n = naggs;
  for (int i = 0; i < naggs; ++i) args.a[i] = aggs[i];
  launch(1, 32, [&] { group_update_serial_kernel(*t, keys, nullptr, n, args, nullptr, error_flag); });
}
}
"""


class GroupTable(C.Structure):
    _fields_ = [("rows", C.c_void_p), ("capacity", C.c_int64), ("row_words", C.c_int32), ("hash_mode", C.c_int32)]


class AggUpdate(C.Structure):
    _fields_ = [("kind", C.c_int32), ("input_type", C.c_int32), ("input", C.c_void_p), ("nulls", C.c_void_p), ("mask", C.c_void_p), ("indices", C.c_void_p),
                ("base_nulls", C.c_void_p), ("acc_word", C.c_int32), ("nonnull_word", C.c_int32)]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    common, agg = source("common.cuh"), source("hash_agg.cu")
    update = between(agg, "__device__ __forceinline__ double input_as_f64", "// ---- keyed hash mode (kHash)")
    update = update.replace("extern __shared__ __align__(16) uint64_t srows[];", "static uint64_t srows[8192];")
    body = BODY % {
        "mixers": between(common, "__host__ __device__ __forceinline__ uint64_t twang_mix64", "__device__ __forceinline__ uint64_t hash_f64"),
        "reductions": between(common, "__device__ __forceinline__ double warp_sum(double v)", "}  // namespace vb2"),
        "update": update,
        "tiny": between(agg, "__device__ __forceinline__ void apply_update_plain", "__global__ void table_init_kernel"),
        "norm": between(agg, "constexpr int kMaxNormCols", "__global__ void minmax_kernel"),
        "occupied": between(agg, "__device__ __forceinline__ bool row_occupied", "__global__ void occupied_bits_kernel"),
        "extract": between(agg, "struct ExtractArgs {", "// Partial results of a fused scan"),
    }
    return build(tmp_path_factory.mktemp("smem_agg_on_host"), "smemagg", body)


def test_block_private_tables_merge_to_the_global_result(host):
    rng = np.random.default_rng(6)
    n, groups, row_words = 4000, 50, 8
    keys = rng.integers(1, groups, n).astype(np.uint64)      # array mode: the normalized key is the slot
    x = np.round(rng.normal(0, 30, n), 3)
    y = rng.integers(-500, 500, n).astype(np.int64)
    yvalid_list = rng.random(n) > 0.1
    yvalid = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    for i in np.nonzero(yvalid_list)[0]:
        yvalid[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    aggs = (AggUpdate * 5)(
        AggUpdate(SUM_F64, DOUBLE, x.ctypes.data, None, None, None, None, 1, -1),
        AggUpdate(SUM_I64, BIGINT, y.ctypes.data, yvalid.ctypes.data, None, None, None, 2, 3),
        AggUpdate(COUNT, BIGINT, None, None, None, None, None, 4, -1),
        AggUpdate(MIN_I64, BIGINT, y.ctypes.data, yvalid.ctypes.data, None, None, None, 5, -1),
        AggUpdate(MAX_F64, DOUBLE, x.ctypes.data, None, None, None, None, 6, -1))

    def run(smem):
        rows = np.zeros(groups * row_words, dtype=np.uint64)
        rows[5::row_words] = np.uint64(np.iinfo(np.int64).max)
        rows[6::row_words] = np.array([-math.inf]).view(np.uint64)[0]
        t = GroupTable(rows.ctypes.data, groups, row_words, 0)
        err = np.zeros(2, dtype=np.int32)
        host.h_group_update(smem, C.byref(t), keys.ctypes.data_as(C.c_void_p), C.c_int64(n), aggs, 5, err.ctypes.data_as(C.c_void_p), 4)
        assert err[0] == 0
        return rows.reshape(groups, row_words)

    via_smem, via_global = run(1), run(0)
    f64 = lambda w: float(np.array([w], dtype=np.uint64).view(np.float64)[0])  # noqa: E731
    i64 = lambda w: int(np.array([w], dtype=np.uint64).view(np.int64)[0])      # noqa: E731
    for g in range(groups):
        sel = keys == g
        if not sel.any():
            assert int(via_smem[g][0]) == 0 and int(via_global[g][0]) == 0
            continue
        ys = y[sel & yvalid_list]
        for table in (via_smem, via_global):
            r = table[g]
            assert int(r[0]) != 0
            assert math.isclose(f64(r[1]), float(x[sel].sum()), rel_tol=1e-12, abs_tol=1e-9)
            assert i64(r[2]) == int(ys.sum()) and int(r[3]) == len(ys) and int(r[4]) == int(sel.sum())
            assert i64(r[5]) == (int(ys.min()) if len(ys) else np.iinfo(np.int64).max)
            assert f64(r[6]) == float(x[sel].max())


def test_tiny_tables_and_serial_batches(host):
    """Up to eight groups (Q1's shape on the general path): register accumulators, one atomic per group and block; very small
    batches: one warp in input order (the reference's sequential accumulation, bit for bit with a Python loop)."""
    rng = np.random.default_rng(9)
    n, groups, row_words = 5000, 6, 4
    keys = rng.integers(0, groups, n).astype(np.uint64)
    x = np.round(rng.normal(0, 30, n), 3)
    y = rng.integers(-500, 500, n).astype(np.int64)
    aggs = (AggUpdate * 3)(AggUpdate(SUM_F64, DOUBLE, x.ctypes.data, None, None, None, None, 1, -1),
                           AggUpdate(SUM_I64, BIGINT, y.ctypes.data, None, None, None, None, 2, -1),
                           AggUpdate(COUNT, BIGINT, None, None, None, None, None, 3, -1))
    f64 = lambda w: float(np.array([w], dtype=np.uint64).view(np.float64)[0])  # noqa: E731
    i64 = lambda w: int(np.array([w], dtype=np.uint64).view(np.int64)[0])      # noqa: E731
    rows = np.zeros(groups * row_words, dtype=np.uint64)
    t = GroupTable(rows.ctypes.data, groups, row_words, 0)
    err = np.zeros(2, dtype=np.int32)
    host.h_group_update_tiny(C.byref(t), keys.ctypes.data_as(C.c_void_p), C.c_int64(n), aggs, 3, err.ctypes.data_as(C.c_void_p), 3)
    assert err[0] == 0
    table = rows.reshape(groups, row_words)
    for g in range(groups):
        sel = keys == g
        assert int(table[g][0]) == int(sel.sum()) == int(table[g][3])  # rows seen, count(*)
        assert math.isclose(f64(table[g][1]), float(x[sel].sum()), rel_tol=1e-12, abs_tol=1e-9) and i64(table[g][2]) == int(y[sel].sum())
    # a 40-row batch through the serial kernel: sums in input order
    m = 40
    rows2 = np.zeros(groups * row_words, dtype=np.uint64)
    t2 = GroupTable(rows2.ctypes.data, groups, row_words, 0)
    host.h_group_update_serial(C.byref(t2), keys.ctypes.data_as(C.c_void_p), C.c_int64(m), aggs, 3, err.ctypes.data_as(C.c_void_p))
    assert err[0] == 0
    table2 = rows2.reshape(groups, row_words)
    for g in range(groups):
        acc, cnt, isum = 0.0, 0, 0
        for i in range(m):
            if keys[i] == g:
                acc += float(x[i])
                isum += int(y[i])
                cnt += 1
        if cnt:
            assert f64(table2[g][1]) == acc and i64(table2[g][2]) == isum and int(table2[g][3]) == cnt  # bit for bit: same order of additions


class ExtractCol(C.Structure):
    _fields_ = [("kind", C.c_int32), ("type", C.c_int32), ("word", C.c_int32), ("count_word", C.c_int32), ("min", C.c_int64), ("mult", C.c_uint64),
                ("range", C.c_uint64), ("null_reserved", C.c_int32), ("pad", C.c_int32), ("values", C.c_void_p), ("valid", C.c_void_p)]


def test_group_by_end_to_end_with_result_extraction(host):
    """normalize -> update -> extract: two nullable keys decoded back from their packed value ids (a NULL key is a group and
    comes back NULL), SUM with its validity from the non-null counter (NULL for a group that saw only NULL inputs,
    SumAggregateBase.h:143-150), AVG = sum / count (AverageAggregateBase.h:86-107), COUNT(*)."""
    from velox_b200.kernels import CColumn
    from velox_b200.vector import INTEGER, flat_vector
    rng = np.random.default_rng(12)
    n = 2000
    k0 = [None if rng.random() < 0.08 else int(v) for v in rng.integers(-4, 5, n)]
    k1 = [None if rng.random() < 0.08 else int(v) for v in rng.integers(100, 106, n)]
    x = [None if rng.random() < 0.5 else float(v) for v in np.round(rng.normal(0, 20, n), 2)]
    key_columns = [flat_vector(BIGINT, k0), flat_vector(INTEGER, k1)]
    cols = (CColumn * 2)(*[c.to_c() for c in key_columns])
    mins = np.array([-4, 100], dtype=np.int64)
    mults = np.array([7, 1], dtype=np.uint64)
    capacity, row_words = 10 * 7 + 2, 4     # packed ids: (0..9) * 7 + (0..6); [rows seen | sum x | non-null x | count(*)]
    rows = np.zeros(capacity * row_words, dtype=np.uint64)
    t = GroupTable(rows.ctypes.data, capacity, row_words, 0)
    xa = np.array([0.0 if v is None else v for v in x])
    xv = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    for i, v in enumerate(x):
        if v is not None:
            xv[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    aggs = (AggUpdate * 2)(AggUpdate(SUM_F64, DOUBLE, xa.ctypes.data, xv.ctypes.data, None, None, None, 1, 2), AggUpdate(COUNT, BIGINT, None, None, None, None, None, 3, -1))
    m = capacity
    out_k0, out_k1 = np.zeros(m, dtype=np.int64), np.zeros(m, dtype=np.int32)
    out_sum, out_avg, out_cnt = np.zeros(m, dtype=np.uint64), np.zeros(m, dtype=np.float64), np.zeros(m, dtype=np.uint64)
    valid = [np.zeros(m // 64 + 2, dtype=np.uint64) for _ in range(4)]
    KEY, WORD, AVG = 1, 2, 4
    outs = (ExtractCol * 5)(
        ExtractCol(KEY, BIGINT, 0, -1, -4, 7, 10, 1, 0, out_k0.ctypes.data, valid[0].ctypes.data),
        ExtractCol(KEY, 3, 0, -1, 100, 1, 7, 1, 0, out_k1.ctypes.data, valid[1].ctypes.data),
        ExtractCol(WORD, DOUBLE, 1, 2, 0, 1, 1, 0, 0, out_sum.ctypes.data, valid[2].ctypes.data),
        ExtractCol(AVG, DOUBLE, 1, 2, 0, 1, 1, 0, 0, out_avg.ctypes.data, valid[3].ctypes.data),
        ExtractCol(WORD, BIGINT, 3, -1, 0, 1, 1, 0, 0, out_cnt.ctypes.data, None))
    scratch = np.zeros(capacity + 8, dtype=np.int32)
    header = np.zeros(2, dtype=np.int64)
    err = np.zeros(2, dtype=np.int32)
    P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    host.h_group_by(cols, 2, P(mins), P(mults), C.c_int64(n), C.byref(t), aggs, 2, outs, 5, P(scratch), P(header), P(err))
    assert err[0] == 0 and int(header[1]) == 0
    want = {}
    for a, b, v in zip(k0, k1, x):
        g = want.setdefault((a, b), [0.0, 0, 0])
        g[2] += 1
        if v is not None:
            g[0] += v
            g[1] += 1
    groups = int(header[0])
    assert groups == len(want)
    bit = lambda words, i: (int(words[i >> 6]) >> (i & 63)) & 1  # noqa: E731
    got = {}
    for i in range(groups):
        a = int(out_k0[i]) if bit(valid[0], i) else None
        b = int(out_k1[i]) if bit(valid[1], i) else None
        s = float(out_sum[i:i + 1].view(np.float64)[0]) if bit(valid[2], i) else None
        avg = float(out_avg[i]) if bit(valid[3], i) else None
        got[(a, b)] = (s, avg, int(out_cnt[i]))
    assert set(got) == set(want)
    for key, (sx, nn, cnt) in want.items():
        s, avg, c = got[key]
        assert c == cnt
        if nn == 0:
            assert s is None and avg is None
        else:
            assert math.isclose(s, sx, rel_tol=1e-12, abs_tol=1e-9) and math.isclose(avg, sx / nn, rel_tol=1e-12, abs_tol=1e-9)
