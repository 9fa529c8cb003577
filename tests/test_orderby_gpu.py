"""OrderBy / TopN on the device against the CPU oracle, ROW BY ROW in order:
both sorts are stable, so the expected sequence is unique. Modelled on the reference's
velox/exec/tests/OrderByTest.cpp and TopNTest.cpp (directions, NULLS FIRST/LAST, multi-key, several
input batches, limits above and below the input size); TPC-H Q1 ends in this operator
(velox/exec/tests/utils/TpchQueryBuilder.cpp:247)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

from oracle import pyoracle
from velox_b200._lib import lib
from velox_b200.plan import PlanBuilder
from velox_b200.task import run_plan
from velox_b200.vector import BIGINT, BOOLEAN, DOUBLE, INTEGER, VARCHAR, dictionary_vector, flat_vector, row_vector

pytestmark = pytest.mark.gpu
NAN = float("nan")


def same_sequence(got, want):
    g, w = got.rows(), want.rows()
    assert len(g) == len(w), (len(g), len(w))
    for i, (a, b) in enumerate(zip(g, w)):
        for x, y in zip(a, b):
            if isinstance(y, float) and math.isnan(y):
                assert isinstance(x, float) and math.isnan(x), (i, a, b)
            else:
                assert x == y and type(x) is type(y), (i, a, b)


def check_ordered(plan, sources, batch_rows=None):
    want = pyoracle.run_plan(plan, sources, threads=1, batch_rows=10000)
    got, stats = run_plan(plan, sources, batch_rows=batch_rows)
    same_sequence(got, want)
    return stats


def table(n, seed, nulls=True):
    rng = np.random.default_rng(seed)
    def maybe(vals, p=0.1):
        return [None if (nulls and rng.random() < p) else v for v in vals]
    x = np.round(rng.normal(0, 5, n), 1)
    x[rng.random(n) < 0.05] = NAN
    x[rng.random(n) < 0.05] = -0.0
    x[rng.random(n) < 0.05] = 0.0
    return row_vector(
        ["id", "k", "big", "x", "b", "s"],
        [flat_vector(BIGINT, np.arange(n)),
         flat_vector(INTEGER, maybe(rng.integers(-5, 6, n).tolist())),
         flat_vector(BIGINT, maybe((rng.integers(-2**40, 2**40, n) * rng.integers(-1000, 1000, n)).tolist())),
         flat_vector(DOUBLE, maybe(x.tolist())),
         flat_vector(BOOLEAN, maybe((rng.random(n) < 0.5).tolist())),
         dictionary_vector(VARCHAR, rng.integers(0, 7, n), ["pear", "apple", "", "fig", "apple pie", "zucchini", None if nulls else "kiwi"])])


KEYSETS = [["k"], ["k DESC"], ["k ASC NULLS FIRST", "x DESC"], ["x", "k DESC NULLS FIRST"], ["s", "b DESC", "big"], ["s DESC NULLS FIRST", "x"],
           ["b", "big DESC"], ["big"], ["x DESC NULLS FIRST"]]


@pytest.mark.parametrize("n,batch", [(1, None), (900, None), (5000, 777), (40000, 9000)])
def test_order_by_matches_oracle(n, batch):
    """n <= 16384 rows: rank-sort kernel; above: LSD radix passes. Ties keep input order on both sides."""
    rv = table(n, seed=n)
    for keys in KEYSETS:
        stats = check_ordered(PlanBuilder().values(rv.names, rv.types).orderBy(keys).planNode(), [rv], batch_rows=batch)
        assert sum(v for k, v in stats.items() if k.endswith("B200OrderBy.b200.sortedRows")) == n


def test_order_by_after_filter_and_dictionary_wraps():
    rv = table(20000, seed=3)
    plan = (PlanBuilder().values(rv.names, rv.types).filter("k is null or k <> 0").project(["id", "x * 2.0 as y", "s", "k"])
            .orderBy(["s", "y DESC", "k"]).planNode())
    check_ordered(plan, [rv], batch_rows=4096)


@pytest.mark.parametrize("count", [1, 10, 5000, 100000])
def test_top_n(count):
    rv = table(30000, seed=9)
    for keys in (["x", "k"], ["s DESC", "big"], ["k DESC NULLS FIRST"]):
        plan = PlanBuilder().values(rv.names, rv.types).topN(keys, count).planNode()
        check_ordered(plan, [rv])
        check_ordered(plan, [rv], batch_rows=4000)


def test_empty_input_and_all_equal_keys():
    rv = table(500, seed=1, nulls=False)
    check_ordered(PlanBuilder().values(rv.names, rv.types).filter("k > 100").orderBy(["k"]).planNode(), [rv])
    const = row_vector(["id", "k"], [flat_vector(BIGINT, np.arange(3000)), flat_vector(INTEGER, np.full(3000, 7, dtype=np.int32))])
    check_ordered(PlanBuilder().values(const.names, const.types).orderBy(["k DESC"]).planNode(), [const], batch_rows=512)  # stable: ids stay ascending


def test_q1_with_its_order_by_on_dbgen_rows():
    """The whole of TPC-H Q1 including the final ORDER BY l_returnflag, l_linestatus, on the dbgen fixture."""
    from test_tpch_reference_data import lineitem_vectors
    fx = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tpch_sf001.npz"))
    li = {k: fx[k] for k in fx.files if k.startswith("l_")}
    rv = lineitem_vectors(li, ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate"])
    plan = (PlanBuilder().values(rv.names, rv.types).filter("l_shipdate <= '1998-09-02'::DATE")
            .project(["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_extendedprice * (1.0 - l_discount) AS d",
                      "l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS c", "l_discount"])
            .partialAggregation(["l_returnflag", "l_linestatus"], ["sum(l_quantity)", "sum(l_extendedprice)", "sum(d)", "sum(c)", "avg(l_quantity)",
                                                                    "avg(l_extendedprice)", "avg(l_discount)", "count(0)"])
            .localPartition([]).finalAggregation().orderBy(["l_returnflag", "l_linestatus"]).planNode())
    want = pyoracle.run_plan(plan, [rv], threads=1, batch_rows=100_000).rows()
    got, _ = run_plan(plan, [rv])
    got = got.rows()
    assert [(r[0], r[1], r[9]) for r in got] == [(r[0], r[1], r[9]) for r in want] == sorted((r[0], r[1], r[9]) for r in want)
    for a, b in zip(got, want):
        for x, y in zip(a[2:9], b[2:9]):
            assert abs(x - y) <= 1e-11 * abs(y)


def test_sort_kernel_large_against_numpy():
    """vb2k_sort_order directly: 3 M rows, (INTEGER desc nulls first, BIGINT asc) — the radix path with a
    null pass — against numpy's stable lexsort."""
    import torch
    from velox_b200.vector import pack_bits
    L = lib()

    class SortKey(C.Structure):
        _fields_ = [("values", C.c_void_p), ("nulls", C.c_void_p), ("type", C.c_int32), ("ascending", C.c_int32), ("nulls_first", C.c_int32),
                    ("significant_bits", C.c_int32)]

    n = 3_000_000
    rng = np.random.default_rng(2)
    a = rng.integers(-1000, 1000, n).astype(np.int32)
    an = rng.random(n) < 0.1
    b = rng.integers(-2**62, 2**62, n)
    da, db = torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda()
    dn = torch.from_numpy(pack_bits(~an).view(np.int64)).cuda()
    keys = (SortKey * 2)(SortKey(da.data_ptr(), dn.data_ptr(), INTEGER, 0, 1, 0), SortKey(db.data_ptr(), None, BIGINT, 1, 0, 0))
    L.vb2k_sort_order_workspace.restype = C.c_size_t
    wsb = L.vb2k_sort_order_workspace(C.c_int64(n), 2)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    order = torch.empty(n, dtype=torch.int32, device="cuda")
    rc = L.vb2k_sort_order(keys, 2, C.c_int64(n), C.c_void_p(order.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_size_t(wsb),
                           C.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0, L.vb2_last_error()
    torch.cuda.synchronize()
    # expected: nulls first, then a descending, ties by b ascending, then input order
    primary = np.where(an, np.int64(-10**9), -a.astype(np.int64))
    want = np.lexsort((b, primary))  # lexsort is stable; last key is the primary
    assert np.array_equal(order.cpu().numpy(), want.astype(np.int32))
