"""High-cardinality aggregation in shared-memory slices (csrc/slice_agg.cu; BASELINE.json configs[4]):
kernel level against numpy (bit exact: integer sums, counts, min / max), operator level against the
oracle, with the slice path forced on small inputs through its thresholds, both partition depths
(256 slices / 256 x P2 slices) and the fall-backs to the table path."""
import ctypes as C

import numpy as np
import pytest

from util import check_plan, stat
from velox_b200._lib import lib
from velox_b200.plan import PlanBuilder
from velox_b200.vector import BIGINT, DOUBLE, flat_vector, row_vector

pytestmark = pytest.mark.gpu

SLICE = {"b200.agg_slice_min_rows": "1000", "b200.agg_slice_min_groups": "100", "b200.fused_pipelines": "false"}
AGG_SUM_F64, AGG_SUM_I64, AGG_COUNT, AGG_MIN_I64, AGG_MAX_I64 = 1, 2, 3, 6, 7
EMPTY = 0xFFFFFFFFFFFFFFFF


class Chunk(C.Structure):
    _fields_ = [("norm_keys", C.c_void_p), ("raw_keys", C.c_void_p), ("key_min", C.c_int64), ("cols", C.c_void_p * 3), ("rows", C.c_int64)]


class Op(C.Structure):
    _fields_ = [("kind", C.c_int32), ("col", C.c_int32), ("word", C.c_int32)]


@pytest.mark.parametrize("n,distinct,estimate", [(3_000_000, 500_000, 500_000), (3_000_000, 500_000, 40_000_000), (200_000, 3_000, 3_000), (5_000, 5_000, 5_000)])
def test_slice_kernels_against_numpy(n, distinct, estimate):
    """estimate = what the HyperLogLog step would hand to level 2: 40 M forces 256 x 128 slices over
    500 K real keys (most slices tiny or empty), the others stay at one level."""
    import torch
    L = lib()
    L.vb2k_slice_agg_workspace.restype = C.c_size_t
    rng = np.random.default_rng(n + distinct)
    keys = rng.integers(0, distinct, n) * 7919 - 3_000_000_000  # raw BIGINT keys, negative too
    v0 = rng.integers(-1000, 1000, n)
    v1 = np.round(rng.normal(0, 10, n), 2)
    kmin = int(keys.min())
    dk, d0, d1 = torch.from_numpy(keys).cuda(), torch.from_numpy(v0).cuda(), torch.from_numpy(v1).cuda()
    half = n // 2 // 64 * 64
    chunks = (Chunk * 2)()
    for i, (a, b) in enumerate(((0, half), (half, n))):
        chunks[i].norm_keys = None
        chunks[i].raw_keys = dk.data_ptr() + a * 8
        chunks[i].key_min = kmin
        chunks[i].cols[0] = d0.data_ptr() + a * 8
        chunks[i].cols[1] = d1.data_ptr() + a * 8
        chunks[i].rows = b - a
    wsb = L.vb2k_slice_agg_workspace(C.c_int64(n), 2)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    regs = (C.c_int32 * L.vb2k_slice_agg_hll_registers())()
    assert L.vb2k_slice_agg_partition(chunks, 2, 2, C.c_int64(n), C.c_void_p(ws.data_ptr()), C.c_size_t(wsb), regs, st) == 0, L.vb2_last_error()
    # group row: [key | sum(v0) | count | max(v0) | sum(v1) | min(v0) | unused | unused]
    ops = (Op * 5)(Op(AGG_SUM_I64, 0, 1), Op(AGG_COUNT, -1, 2), Op(AGG_MAX_I64, 0, 3), Op(AGG_SUM_F64, 1, 4), Op(AGG_MIN_I64, 0, 5))
    rw = 8
    init = np.zeros(rw, dtype=np.uint64)
    init[3] = np.uint64(np.int64(-2**63).astype(np.uint64))
    init[5] = np.uint64(2**63 - 1)
    init[6] = np.uint64(12345)
    L.vb2k_slice_agg_output_rows.restype = C.c_int64
    cap = L.vb2k_slice_agg_output_rows(C.c_int64(estimate))
    assert cap >= distinct
    # table-shaped output: every row starts as [EMPTY | row_init[1:]]; blocks fill the rows of the chunks they reserve
    rows = torch.from_numpy(np.tile(np.concatenate([[np.uint64(EMPTY)], init[1:]]), cap).view(np.int64)).cuda()
    words = torch.zeros(4, dtype=torch.int64, device="cuda")  # [0] groups, [1] rows reserved, [2] error, [3] overflow
    rc = L.vb2k_slice_agg_finish(C.c_int64(n), 2, C.c_int64(estimate), ops, 5, rw, init.ctypes.data_as(C.c_void_p), C.c_void_p(rows.data_ptr()), C.c_int64(cap),
                                 C.c_void_p(words.data_ptr()), C.c_void_p(words.data_ptr() + 8), C.c_void_p(words.data_ptr() + 16), C.c_void_p(words.data_ptr() + 24),
                                 C.c_void_p(ws.data_ptr()), C.c_size_t(wsb), st)
    assert rc == 0, L.vb2_last_error()
    torch.cuda.synchronize()
    w = words.cpu().numpy()
    uk, inv = np.unique(keys, return_inverse=True)
    assert w[2] == 0 and w[3] == 0 and w[0] == len(uk) and len(uk) <= w[1] <= cap, w
    got = rows.cpu().numpy().reshape(cap, rw)
    got = got[got[:, 0].view(np.uint64) != np.uint64(EMPTY)]  # occupied rows (the unused tails of the chunks stay EMPTY)
    assert len(got) == len(uk)
    got = got[np.argsort(got[:, 0])]
    assert np.array_equal(got[:, 0], uk - kmin + 1)                       # normalized keys, each exactly once
    order = np.argsort(inv, kind="stable")
    starts = np.searchsorted(inv[order], np.arange(len(uk)))
    assert np.array_equal(got[:, 1], np.add.reduceat(v0[order], starts))
    assert np.array_equal(got[:, 2], np.bincount(inv, minlength=len(uk)))
    assert np.array_equal(got[:, 3], np.maximum.reduceat(v0[order], starts))
    assert np.array_equal(got[:, 5], np.minimum.reduceat(v0[order], starts))
    want1 = np.add.reduceat(v1[order], starts)
    assert np.allclose(got[:, 4].view(np.float64), want1, rtol=1e-12, atol=1e-9)
    assert np.all(got[:, 6] == 12345) and np.all(got[:, 7] == 0)         # untouched words come from row_init


def table(n, distinct, seed=0):
    rng = np.random.default_rng(seed)
    return row_vector(["k", "v", "x"], [flat_vector(BIGINT, rng.integers(0, distinct, n) * 104729 - 10**12),
                                        flat_vector(BIGINT, rng.integers(-500, 500, n)),
                                        flat_vector(DOUBLE, np.round(rng.normal(0, 100, n), 3))])


@pytest.mark.parametrize("hint", [None, "30000000"])
def test_operator_takes_the_slice_path(hint):
    rv = table(400_000, 150_000)
    cfg = dict(SLICE)
    if hint:
        cfg["b200.agg_slice_distinct_hint"] = hint  # sizes level 2 as if there were 30 M groups: 256 x 64 slices
    aggs = ["sum(v)", "count(0)", "min(v)", "max(x)", "avg(x)", "sum(x)"]
    single = PlanBuilder().values(rv.names, rv.types).singleAggregation(["k"], aggs).planNode()
    (st,) = check_plan(single, [rv], configs=(cfg,), rel_tol=1e-11)
    assert stat(st, "b200.sliceAggRows") == 400_000 and stat(st, "b200.sliceAggGroups") > 100_000
    (st,) = check_plan(single, [rv], configs=(cfg,), batch_rows=64 * 1024, rel_tol=1e-11)  # several buffered batches
    assert stat(st, "b200.sliceAggRows") == 400_000
    two = PlanBuilder().values(rv.names, rv.types).partialAggregation(["k"], aggs).localPartition([]).finalAggregation().planNode()
    check_plan(two, [rv], configs=(dict(cfg, **{"b200.fused_pipelines": "false"}),), rel_tol=1e-11)


def test_slice_path_falls_back_to_the_table():
    """Low cardinality stays on the table path; NULL inputs arriving after buffering began drain the
    buffer through the table path; a table too small for the keys of a slice is detected."""
    rv = table(300_000, 50)
    plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(["k"], ["sum(v)", "count(0)"]).planNode()
    (st,) = check_plan(plan, [rv], configs=(SLICE,))
    assert stat(st, "b200.sliceAggRows") == 0
    n = 200_000
    rng = np.random.default_rng(4)
    vals = [None if (i >= n // 2 and rng.random() < 0.2) else int(i % 997) for i in range(n)]  # NULLs only in the second half
    rvn = row_vector(["k", "v"], [flat_vector(BIGINT, rng.integers(0, 80_000, n)), flat_vector(BIGINT, vals)])
    plan = PlanBuilder().values(rvn.names, rvn.types).singleAggregation(["k"], ["sum(v)", "count(v)", "count(0)"]).planNode()
    (st,) = check_plan(plan, [rvn], configs=(SLICE,), batch_rows=50_000)
    assert stat(st, "b200.sliceAggRows") == 0 and stat(st, "b200.genericBatches") >= 4
    # distinct hint far too low: the output array planned from it fills up, the operator redoes the input on the table path
    rv = table(400_000, 300_000, seed=2)
    plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(["k"], ["sum(v)", "count(0)"]).planNode()
    (st,) = check_plan(plan, [rv], configs=(dict(SLICE, **{"b200.agg_slice_distinct_hint": "2000"}),))
    assert stat(st, "b200.sliceAggRows") == 0 and stat(st, "b200.genericBatches") >= 1


def test_slice_path_wide_integer_sums_and_overflow():
    """BIGINT inputs outside the int32 range take the 64-bit shared-memory add (values inside it are carried as two 32-bit
    halves): mixed magnitudes and signs sum exactly; a sum that leaves the int64 range raises the reference's user error
    (functions/prestosql/aggregates/SumAggregate.cpp: checked addition)."""
    from util import check_user_error
    n, distinct = 300_000, 60_000
    rng = np.random.default_rng(9)
    mags = rng.choice([1, 1000, 2**31 - 1, 2**31, 2**40, 2**45], n)
    v = (rng.integers(0, 1000, n) + 1) * mags * rng.choice([-1, 1], n)
    rv = row_vector(["k", "v"], [flat_vector(BIGINT, rng.integers(0, distinct, n) * 31 - 5_000_000), flat_vector(BIGINT, v)])
    plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(["k"], ["sum(v)", "count(0)", "min(v)", "max(v)"]).planNode()
    (st,) = check_plan(plan, [rv], configs=(SLICE,))
    assert stat(st, "b200.sliceAggRows") == n
    big = np.full(n, 2**62, dtype=np.int64)
    big[: n // 2] = 1
    ro = row_vector(["k", "v"], [flat_vector(BIGINT, rng.integers(0, distinct, n)), flat_vector(BIGINT, big)])
    check_user_error(PlanBuilder().values(ro.names, ro.types).singleAggregation(["k"], ["sum(v)"]).planNode(), [ro], configs=(SLICE,))
