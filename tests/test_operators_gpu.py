"""Operator-level parity: plans run through the C ABI (Task -> Driver -> B200 operators installed
by the DriverAdapter) against the CPU oracle. Modelled on the reference's
velox/exec/tests/FilterProjectTest.cpp, AggregationTest.cpp, HashJoinTest.cpp and the scalar
known-answer tests in velox/functions/prestosql/tests/{Arithmetic,Comparisons}Test.cpp."""
import numpy as np
import pytest

from velox_b200 import tpch
from velox_b200.plan import PlanBuilder
from velox_b200.vector import (BIGINT, BOOLEAN, DOUBLE, INTEGER, VARCHAR, constant_vector, dictionary_vector, flat_vector,
                               row_vector)

from util import FUSED, GENERIC, check_plan, check_user_error, stat

pytestmark = pytest.mark.gpu


@pytest.fixture(params=["jit", "interpreter"], autouse=True)
def expression_engine(request):
    """Every plan runs with both expression engines: NVRTC-compiled kernels and the interpreter."""
    from velox_b200._lib import lib
    lib().vb2k_set_expression_jit(1 if request.param == "jit" else 0)
    yield request.param
    lib().vb2k_set_expression_jit(1)

NAN, INF = float("nan"), float("inf")


def table(n=1000, seed=0, nulls=True):
    rng = np.random.default_rng(seed)
    def maybe(vals, p=0.1):
        if not nulls:
            return list(vals)
        return [None if rng.random() < p else v for v in vals]
    return row_vector(
        ["c0", "c1", "c2", "c3", "c4", "c5"],
        [flat_vector(BIGINT, maybe(rng.integers(-50, 50, n).tolist())),
         flat_vector(INTEGER, maybe(rng.integers(0, 10, n).tolist())),
         flat_vector(DOUBLE, maybe(np.round(rng.normal(0, 100, n), 3).tolist())),
         flat_vector(DOUBLE, maybe(rng.integers(0, 20, n).astype(float).tolist())),
         flat_vector(BOOLEAN, maybe((rng.random(n) < 0.5).tolist())),
         dictionary_vector(VARCHAR, rng.integers(0, 5, n), ["apple", "banana", "cherry", "date", None if nulls else "egg"])])


# ---- FilterProject / expression engine --------------------------------------------------------
@pytest.mark.parametrize("nulls", [False, True])
def test_filter(nulls):
    rv = table(nulls=nulls)
    for f in ["c0 < 10", "c2 >= 0.0 and c3 < 10.0", "c0 % 2 = 0 or c1 > 7", "c4", "not c4", "c2 is null", "c2 is not null",
              "c0 between -10 and 10", "c5 = 'banana'", "c5 like 'b%'", "c5 like '%e%' and c1 <> 3", "c3 not between 5.0 and 9.0"]:
        check_plan(PlanBuilder().values(rv.names, rv.types).filter(f).planNode(), [rv])


@pytest.mark.parametrize("nulls", [False, True])
def test_project(nulls):
    rv = table(nulls=nulls)
    check_plan(PlanBuilder().values(rv.names, rv.types).project(
        ["c0", "c0 + 1", "c2 * (1.0 - c3)", "c2 / c3", "c1 - 4", "cast(c1 as bigint) * c0", "cast(c0 as double) + c2",
         "case when c0 > 0 then c2 else 0.0 end", "case when c4 then c0 when c1 > 5 then c0 * 2 end", "c2 < c3", "c5",
         "c0 is null or c1 is null", "-c2", "c4 and c0 > 0", "c4 or c0 > 0", "if(c1 = 3, c1, c1 + 100)"]).planNode(), [rv])


def test_filter_project_all_filtered_and_all_pass():
    rv = table(nulls=False)
    check_plan(PlanBuilder().values(rv.names, rv.types).filter("c0 > 1000").project(["c0", "c2 * 2.0"]).planNode(), [rv])
    check_plan(PlanBuilder().values(rv.names, rv.types).filter("c0 > -1000").project(["c0", "c2 * 2.0", "c5"]).planNode(), [rv])


def test_filter_project_encodings_and_batches():
    """Dictionary / constant inputs, dictionary over the filter's own wrap (two filters), several batches."""
    n = 5000
    rng = np.random.default_rng(3)
    rv = row_vector(["a", "b", "c", "d"], [
        dictionary_vector(DOUBLE, rng.integers(0, 7, n), [1.5, -2.0, None, 4.25, NAN, 0.0, 1e10], index_nulls=rng.random(n) < 0.05),
        constant_vector(BIGINT, 7, n),
        dictionary_vector(BIGINT, rng.integers(0, 3, n), [10, 20, 30]),
        constant_vector(DOUBLE, None, n)])
    plan = (PlanBuilder().values(rv.names, rv.types).filter("c > 10").project(["a", "b", "c", "d", "a * 2.0 as e"])
            .filter("b = 7 and (e > 0.0 or e is null)").project(["a", "c", "e", "d is null as f", "cast(c as double) + e"]).planNode())
    check_plan(plan, [rv])
    check_plan(plan, [rv], batch_rows=1024)


def test_nan_comparisons():
    """NaN is the largest value and equal to itself (ComparisonsTest.cpp:650-720)."""
    vals = [1.0, NAN, -INF, INF, 0.0, -0.0, NAN, None]
    rv = row_vector(["a", "b"], [flat_vector(DOUBLE, vals), flat_vector(DOUBLE, vals[::-1])])
    check_plan(PlanBuilder().values(rv.names, rv.types).project(
        ["a < b", "a <= b", "a > b", "a >= b", "a = b", "a <> b", "a between 0.0 and b", "a < 1.0", "a >= 1.0"]).planNode(), [rv])


def test_three_valued_logic():
    t, f, n = True, False, None
    a = [t, t, t, f, f, f, n, n, n]
    b = [t, f, n, t, f, n, t, f, n]
    rv = row_vector(["a", "b"], [flat_vector(BOOLEAN, a), flat_vector(BOOLEAN, b)])
    check_plan(PlanBuilder().values(rv.names, rv.types).project(["a and b", "a or b", "not a", "a and b and a", "a or b or a"]).planNode(), [rv])
    check_plan(PlanBuilder().values(rv.names, rv.types).filter("a or b").planNode(), [rv])


def test_checked_arithmetic_errors():
    big = 2**62
    rv = row_vector(["a", "b"], [flat_vector(BIGINT, [1, big, -big, 5]), flat_vector(BIGINT, [2, big, -big - 5, 0])])
    ok = PlanBuilder().values(rv.names, rv.types).project(["a + 1", "a - b"]).planNode()
    check_plan(ok, [row_vector(["a", "b"], [flat_vector(BIGINT, [1, 2]), flat_vector(BIGINT, [3, 4])])])
    for expr in ["a + b", "a * b", "a / b", "a % b", "a - b - b - b"]:
        check_user_error(PlanBuilder().values(rv.names, rv.types).project([expr]).planNode(), [rv])
    # errors on rows that end up not selected are suppressed inside AND / CASE (ConjunctExpr.cpp:98-99)
    check_plan(PlanBuilder().values(rv.names, rv.types).filter("b <> 0 and a / b > 0").planNode(), [rv])
    div = row_vector(["a", "b"], [flat_vector(BIGINT, [6, 5, 7, None]), flat_vector(BIGINT, [3, 0, 0, 0])])
    for e in ["a / b > 0 and b <> 0", "b <> 0 and a / b > 0", "a / b > 0 or b = 0"]:  # FALSE / TRUE dominate the error in either order
        check_plan(PlanBuilder().values(div.names, div.types).project([e]).planNode(), [div])
    check_user_error(PlanBuilder().values(div.names, div.types).project(["a / b > 0 and a > 0"]).planNode(), [div])
    check_plan(PlanBuilder().values(rv.names, rv.types).project(["case when b <> 0 then a / b else 0 end"]).planNode(),
               [row_vector(["a", "b"], [flat_vector(BIGINT, [6, 5]), flat_vector(BIGINT, [3, 0])])])
    i32 = row_vector(["a"], [flat_vector(INTEGER, [2**31 - 1, 1])])
    check_user_error(PlanBuilder().values(i32.names, i32.types).project(["a + a"]).planNode(), [i32])
    dbl = row_vector(["a"], [flat_vector(DOUBLE, [1.5, NAN])])
    check_user_error(PlanBuilder().values(dbl.names, dbl.types).project(["cast(a as bigint)"]).planNode(), [dbl])


def test_flat_varchar_keys_group_sort_and_filter():
    """Flat (non-dictionary) VARCHAR columns — TPC-H's l_returnflag / l_linestatus are flat VARCHAR(1) in the
    reference's generator (tpch/gen/TpchGen.cpp:278-317) — as grouping keys, sort keys, filter operands and
    pass-through outputs: they are dictionary-encoded on upload (the device analogue of VectorHasher's
    short-strings-as-numbers, exec/VectorHasher.h:377-387)."""
    rng = np.random.default_rng(21)
    n = 6000
    flags = [None if rng.random() < 0.05 else "ANR"[i] for i in rng.integers(0, 3, n)]
    names = [None if rng.random() < 0.05 else f"name-{i:04d}-with-more-than-twelve-bytes" for i in rng.integers(0, 700, n)]
    rv = row_vector(["f", "s", "v", "k"], [flat_vector(VARCHAR, flags), flat_vector(VARCHAR, names), flat_vector(DOUBLE, np.round(rng.normal(0, 10, n), 2)),
                                           flat_vector(BIGINT, rng.integers(0, 5, n))])
    for cfg in (FUSED, GENERIC):
        check_plan(PlanBuilder().values(rv.names, rv.types).singleAggregation(["f"], ["sum(v)", "count(0)", "max(k)"]).planNode(), [rv], configs=(cfg,))
        check_plan(PlanBuilder().values(rv.names, rv.types).partialAggregation(["s", "k"], ["sum(v)", "avg(v)"]).localPartition([]).finalAggregation().planNode(),
                   [rv], configs=(cfg,), batch_rows=1500, rel_tol=1e-11)  # several batches, each with its own dictionary
    check_plan(PlanBuilder().values(rv.names, rv.types).filter("f = 'N' or s like 'name-00%'").project(["s", "f", "v * 2.0 as w"]).planNode(), [rv], batch_rows=2000)
    from oracle import pyoracle
    from velox_b200.task import run_plan
    plan = PlanBuilder().values(rv.names, rv.types).project(["s", "f", "k"]).orderBy(["f DESC NULLS FIRST", "s", "k"]).planNode()
    want = pyoracle.run_plan(plan, [rv], threads=1).rows()
    got, _ = run_plan(plan, [rv], batch_rows=2500)
    assert [r[:2] for r in got.rows()] == [r[:2] for r in want]


def test_cast_to_boolean():
    """cast(x as boolean) is x != 0, NaN included (velox/type/Conversions.h:158-207, folly::to<bool>); the
    result is a normalised 0/1 byte, usable as a filter."""
    rv = row_vector(["a", "b", "c"], [flat_vector(BIGINT, [5, 0, -3, None, 256]), flat_vector(INTEGER, [0, 7, None, 1, 65536]),
                                        flat_vector(DOUBLE, [0.4, 0.0, NAN, None, -0.0])])
    check_plan(PlanBuilder().values(rv.names, rv.types).project(["cast(a as boolean)", "cast(b as boolean)", "cast(c as boolean)",
                                                                  "cast(a as boolean) and cast(c as boolean)"]).planNode(), [rv])
    check_plan(PlanBuilder().values(rv.names, rv.types).filter("cast(a as boolean)").planNode(), [rv])


def test_plumbing_config_exprset():
    """BASELINE.json configs[0]: l_extendedprice*(1-l_discount) WHERE l_quantity<24 on 1M-row
    DOUBLE/BIGINT flat vectors."""
    n = 1_000_000
    t = tpch.gen_lineitem(n, 1000, seed=21, device="cpu")
    rv = row_vector(["l_partkey", "l_quantity", "l_extendedprice", "l_discount"],
                    [flat_vector(BIGINT, t["l_partkey"].numpy()), flat_vector(DOUBLE, t["l_quantity"].numpy()),
                     flat_vector(DOUBLE, t["l_extendedprice"].numpy()), flat_vector(DOUBLE, t["l_discount"].numpy())])
    plan = (PlanBuilder().values(rv.names, rv.types).filter("l_quantity < 24.0")
            .project(["l_partkey", "l_extendedprice * (1.0 - l_discount)"]).planNode())
    check_plan(plan, [rv], oracle_batch_rows=100_000)


# ---- HashAggregation ------------------------------------------------------------------------------
AGGS = ["sum(c0)", "sum(c1)", "sum(c2)", "avg(c2)", "avg(c0)", "count(0)", "count(c2)", "min(c0)", "max(c2)", "min(c1)", "max(c0)"]


@pytest.mark.parametrize("keys", [[], ["c1"], ["c5"], ["c1", "c5"], ["c4", "c1", "c5"], ["c0"]])
@pytest.mark.parametrize("nulls", [False, True])
def test_aggregation_single(keys, nulls):
    rv = table(n=3000, seed=5, nulls=nulls)
    plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(keys, AGGS).planNode()
    check_plan(plan, [rv], rel_tol=1e-11)
    check_plan(plan, [rv], batch_rows=700, rel_tol=1e-11)


@pytest.mark.parametrize("keys", [[], ["c1", "c5"]])
def test_aggregation_partial_final(keys):
    rv = table(n=3000, seed=6)
    plan = PlanBuilder().values(rv.names, rv.types).partialAggregation(keys, AGGS).localPartition([]).finalAggregation().planNode()
    check_plan(plan, [rv], rel_tol=1e-11)
    plan = (PlanBuilder().values(rv.names, rv.types).partialAggregation(keys, AGGS).intermediateAggregation()
            .finalAggregation().planNode())
    check_plan(plan, [rv], batch_rows=900, rel_tol=1e-11)
    # partial output itself (intermediate types)
    check_plan(PlanBuilder().values(rv.names, rv.types).partialAggregation(keys, AGGS).planNode(), [rv], rel_tol=1e-11)


def test_aggregation_empty_and_all_null():
    rv = table(n=50, seed=1)
    # empty input after a filter: global aggregation still emits one row (sum NULL, count 0)
    plan = PlanBuilder().values(rv.names, rv.types).filter("c0 > 1000").singleAggregation([], ["sum(c2)", "count(0)", "avg(c0)", "min(c1)"]).planNode()
    check_plan(plan, [rv])
    plan = PlanBuilder().values(rv.names, rv.types).filter("c0 > 1000").singleAggregation(["c1"], ["sum(c2)", "count(0)"]).planNode()
    check_plan(plan, [rv])
    allnull = row_vector(["k", "v"], [flat_vector(INTEGER, [1, 1, 2]), flat_vector(DOUBLE, [None, None, None])])
    check_plan(PlanBuilder().values(allnull.names, allnull.types).singleAggregation(["k"], ["sum(v)", "avg(v)", "count(v)", "max(v)"]).planNode(), [allnull])


def test_aggregation_nulls_appear_in_a_later_batch():
    """The non-null counters of sum / min / max start being tracked when the first NULL (or mask)
    shows up: groups seen only in earlier all-non-null batches must stay non-NULL, groups whose
    inputs are all NULL must come out NULL (SumAggregateBase.h:71-142 null handling)."""
    k = [1, 2, 3, 1, 2, 3] + [3, 4, 5, 4, 5, 5]
    v = [1.0, 2.0, 3.0, 4.0, 5.0, 6.0] + [None, None, 7.0, None, 8.0, None]
    w = [10, 20, 30, 40, 50, 60] + [None, None, None, 1, None, None]
    rv = row_vector(["k", "v", "w"], [flat_vector(INTEGER, k), flat_vector(DOUBLE, v), flat_vector(BIGINT, w)])
    aggs = ["sum(v)", "min(v)", "max(w)", "sum(w)", "avg(v)", "count(v)", "count(0)"]
    for keys in (["k"], []):
        plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(keys, aggs).planNode()
        check_plan(plan, [rv], batch_rows=6)   # batch 1 has no NULLs, batch 2 does
        check_plan(plan, [rv], batch_rows=12)
    # hash mode (key range beyond array mode) with the same late NULLs
    big = row_vector(["k", "v", "w"], [flat_vector(BIGINT, [x * (1 << 40) for x in k]), flat_vector(DOUBLE, v), flat_vector(BIGINT, w)])
    (st,) = check_plan(PlanBuilder().values(big.names, big.types).singleAggregation(["k"], aggs).planNode(), [big], batch_rows=6)
    assert stat(st, "b200.aggMode") == 2


def test_aggregation_many_aggregates():
    """More aggregate updates than one vb2k_group_update call carries (16): later slices find the groups again."""
    rv = table(n=3000, seed=21)
    aggs = [f"{fn}({c})" for fn in ("sum", "min", "max", "avg", "count") for c in ("c0", "c2", "c3")] + ["count(0)", "sum(c1)", "max(c1)"]
    assert len(aggs) > 16
    for keys in (["c1"], ["c0", "c5"], []):
        check_plan(PlanBuilder().values(rv.names, rv.types).singleAggregation(keys, aggs).planNode(), [rv], batch_rows=700, rel_tol=1e-11)


def test_aggregation_sliced_batches_all_encodings():
    """A batch larger than the probe chunk is consumed as zero-copy slices: every encoding (flat,
    dictionary VARCHAR, packed BOOLEAN, validity bitmaps) must slice correctly."""
    rv = table(n=200_000, seed=33)
    aggs = ["sum(c2)", "count(0)", "max(c3)", "count(c4)", "min(c0)"]
    for keys in (["c1"], ["c5", "c4"], ["c0", "c1"]):
        plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(keys, aggs).planNode()
        check_plan(plan, [rv], configs=({"b200.agg_probe_chunk_rows": "65536"},), rel_tol=1e-10, oracle_batch_rows=50_000)


def test_aggregation_masks():
    rv = table(n=2000, seed=9)
    plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(["c1"], ["sum(c2)", "count(0)", "avg(c0)"], masks=["c4", "c4", None]).planNode()
    check_plan(plan, [rv], rel_tol=1e-11)


def test_aggregation_sum_overflow():
    rv = row_vector(["k", "v"], [flat_vector(INTEGER, [1, 1, 2]), flat_vector(BIGINT, [2**62, 2**62, 5])])
    check_user_error(PlanBuilder().values(rv.names, rv.types).singleAggregation(["k"], ["sum(v)"]).planNode(), [rv])
    check_user_error(PlanBuilder().values(rv.names, rv.types).singleAggregation([], ["sum(v)"]).planNode(), [rv])


def test_aggregation_high_cardinality_hash_mode():
    """BASELINE.json configs[4] in miniature: BIGINT keys whose range exceeds array mode."""
    n = 400_000
    rng = np.random.default_rng(11)
    keys = rng.integers(0, 2**40, 60_000)[rng.integers(0, 60_000, n)]
    rv = row_vector(["k", "v"], [flat_vector(BIGINT, keys), flat_vector(BIGINT, (np.arange(n) % 1000))])
    plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(["k"], ["sum(v)", "count(0)", "min(v)"]).planNode()
    (st,) = check_plan(plan, [rv], oracle_batch_rows=100_000)
    assert stat(st, "b200.aggMode") == 2  # hash mode
    check_plan(plan, [rv], batch_rows=50_000, oracle_batch_rows=100_000)  # growth / rehash across batches
    # bounded find-or-insert passes inside one batch (the table grows between passes)
    (st,) = check_plan(plan, [rv], configs=({"b200.agg_probe_chunk_rows": "65536"},), oracle_batch_rows=100_000)
    assert stat(st, "b200.genericBatches") == 7 and stat(st, "b200.aggRelayouts") >= 1
    two = row_vector(["a", "b", "v"], [flat_vector(BIGINT, keys), flat_vector(INTEGER, (keys % 7).astype(np.int32)), flat_vector(DOUBLE, np.ones(n))])
    check_plan(PlanBuilder().values(two.names, two.types).singleAggregation(["a", "b"], ["sum(v)", "count(0)"]).planNode(), [two], oracle_batch_rows=100_000)


def _lineitem(n, seed, nparts=2000):
    t = tpch.gen_lineitem(n, nparts, seed=seed, device="cpu")
    h = {k: v.numpy() for k, v in t.items()}
    def col(name):
        if name == "l_returnflag":
            return dictionary_vector(VARCHAR, h[name], tpch.RETURNFLAG_DICT)
        if name == "l_linestatus":
            return dictionary_vector(VARCHAR, h[name], tpch.LINESTATUS_DICT)
        if name == "l_shipdate":
            return flat_vector(INTEGER, h[name])
        if name == "l_partkey":
            return flat_vector(BIGINT, h[name])
        return flat_vector(DOUBLE, h[name])
    return h, col


def q1_plan(rv):
    return (PlanBuilder().values(rv.names, rv.types)
            .filter("l_shipdate < '1998-09-03'::DATE")
            .project(["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice",
                      "l_extendedprice * (1.0 - l_discount) AS l_sum_disc_price",
                      "l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS l_sum_charge", "l_discount"])
            .partialAggregation(["l_returnflag", "l_linestatus"],
                                ["sum(l_quantity)", "sum(l_extendedprice)", "sum(l_sum_disc_price)", "sum(l_sum_charge)",
                                 "avg(l_quantity)", "avg(l_extendedprice)", "avg(l_discount)", "count(0)"])
            .localPartition([]).finalAggregation().planNode())


def q6_plan(rv):
    return (PlanBuilder().values(rv.names, rv.types)
            .filter("l_shipdate between '1994-01-01'::DATE and '1994-12-31'::DATE and "
                    "l_discount between 0.05 and 0.07 and l_quantity < 24.0")
            .project(["l_extendedprice * l_discount"])
            .partialAggregation([], ["sum(p0)"]).localPartition([]).finalAggregation().planNode())


def q14_plan(li, pt):
    build = PlanBuilder().values(pt.names, pt.types, source=1)
    return (PlanBuilder().values(li.names, li.types, source=0)
            .filter("l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE")
            .project(["l_extendedprice * (1.0 - l_discount) as part_revenue", "l_shipdate", "l_partkey"])
            .hashJoin(["l_partkey"], ["p_partkey"], build, "", ["part_revenue", "p_type"])
            .project(["(CASE WHEN (p_type LIKE 'PROMO%') THEN part_revenue ELSE 0.0 END) as filter_revenue", "part_revenue"])
            .partialAggregation([], ["sum(part_revenue) as total_revenue", "sum(filter_revenue) as total_promo_revenue"])
            .localPartition([]).finalAggregation()
            .project(["100.00 * total_promo_revenue/total_revenue as promo_revenue"]).planNode())


@pytest.mark.parametrize("n", [1000, 250_000])
def test_tpch_q1_q6_fused_and_generic(n):
    """TPC-H Q1 / Q6 (TpchQueryBuilder.cpp:203-256,756-788): the fused kernel and the generic
    operator chain both match the oracle; the fused path must really have been taken."""
    h, col = _lineitem(n, seed=31)
    names = ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate"]
    rv = row_vector(names, [col(c) for c in names])
    tol = max(1e-12, n * 2.0 ** -53)
    st_f, st_g = check_plan(q1_plan(rv), [rv], configs=(FUSED, GENERIC), rel_tol=tol, oracle_batch_rows=100_000)
    # partial + final collapse into one aggregation (adapter) that takes the fused kernel; nothing is merged generically
    assert stat(st_f, "b200.fusedBatches") == 1 and stat(st_f, "b200.genericBatches") == 0
    assert stat(st_g, "b200.fusedBatches") == 0 and stat(st_g, "b200.genericBatches") == 2
    check_plan(q1_plan(rv), [rv], configs=(FUSED, GENERIC), batch_rows=60_000, rel_tol=tol, oracle_batch_rows=100_000)
    names6 = ["l_shipdate", "l_extendedprice", "l_quantity", "l_discount"]
    rv6 = row_vector(names6, [col(c) for c in names6])
    st_f, st_g = check_plan(q6_plan(rv6), [rv6], configs=(FUSED, GENERIC), rel_tol=tol, oracle_batch_rows=100_000)
    assert stat(st_f, "b200.fusedBatches") == 1 and stat(st_g, "b200.fusedBatches") == 0


def test_tpch_q1_nulls_fall_back_to_generic_kernels():
    """A batch with NULLs cannot take the fused kernel; the same operator then runs the generic
    kernels for it and the result still matches."""
    h, col = _lineitem(5000, seed=2)
    names = ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate"]
    cols = [col(c) for c in names]
    qn = np.zeros(5000, dtype=bool)
    qn[::17] = True
    cols[2] = flat_vector(DOUBLE, h["l_quantity"], qn)
    rv = row_vector(names, cols)
    (st,) = check_plan(q1_plan(rv), [rv], rel_tol=1e-12)
    assert stat(st, "b200.fusedBatches") == 0 and stat(st, "b200.genericBatches") == 1  # partial + final collapsed into one aggregation


# ---- HashBuild / HashProbe ----------------------------------------------------------------------
def _join_tables(seed=0, nulls=True, dup=True):
    rng = np.random.default_rng(seed)
    n, m = 2000, 300
    pk = rng.integers(0, 400, n).tolist()
    bk = (rng.integers(0, 350, m) if dup else rng.permutation(400)[:m]).tolist()
    if nulls:
        pk = [None if rng.random() < 0.05 else v for v in pk]
        bk = [None if rng.random() < 0.05 else v for v in bk]
    probe = row_vector(["pk", "pv", "ps"], [flat_vector(BIGINT, pk), flat_vector(DOUBLE, rng.normal(size=n).round(3).tolist()),
                                             dictionary_vector(VARCHAR, rng.integers(0, 3, n), ["x", "y", "z"])])
    build = row_vector(["bk", "bv", "bs"], [flat_vector(BIGINT, bk), flat_vector(INTEGER, rng.integers(0, 100, m).tolist()),
                                             dictionary_vector(VARCHAR, rng.integers(0, 4, m), ["red", "green", "blue", None])])
    return probe, build


@pytest.mark.parametrize("join_type", ["inner", "left", "semi", "anti"])
@pytest.mark.parametrize("dup", [False, True])
def test_hash_join_types(join_type, dup):
    probe, build = _join_tables(seed=4, dup=dup)
    b = PlanBuilder().values(build.names, build.types, source=1)
    outs = ["pk", "pv", "ps"] if join_type in ("semi", "anti") else ["pk", "pv", "ps", "bv", "bs"]
    plan = PlanBuilder().values(probe.names, probe.types, source=0).hashJoin(["pk"], ["bk"], b, "", outs, joinType=join_type).planNode()
    (st,) = check_plan(plan, [probe, build])
    # unique build keys take the one-pass probe (warp-ballot match bitmap), duplicates the count / scan / emit chain walk
    assert (stat(st, "b200.uniqueKeyProbes") > 0) == (not dup)
    check_plan(plan, [probe, build], batch_rows=512)


def test_hash_join_filter_multikey_and_empty_build():
    probe, build = _join_tables(seed=8)
    b = PlanBuilder().values(build.names, build.types, source=1)
    plan = PlanBuilder().values(probe.names, probe.types, source=0).hashJoin(["pk"], ["bk"], b, "pv > 0.0 and bv < 50", ["pk", "pv", "bv"]).planNode()
    check_plan(plan, [probe, build])
    # two keys (BIGINT + INTEGER)
    rng = np.random.default_rng(1)
    p2 = row_vector(["a", "b", "v"], [flat_vector(BIGINT, rng.integers(0, 20, 1000)), flat_vector(INTEGER, rng.integers(0, 5, 1000).astype(np.int32)),
                                      flat_vector(DOUBLE, rng.normal(size=1000))])
    b2 = row_vector(["x", "y", "w"], [flat_vector(BIGINT, rng.integers(0, 25, 80)), flat_vector(INTEGER, rng.integers(0, 6, 80).astype(np.int32)),
                                      flat_vector(BIGINT, rng.integers(0, 1000, 80))])
    bb = PlanBuilder().values(b2.names, b2.types, source=1)
    check_plan(PlanBuilder().values(p2.names, p2.types, source=0).hashJoin(["a", "b"], ["x", "y"], bb, "", ["a", "b", "v", "w"]).planNode(), [p2, b2])
    # sparse keys -> hash-mode table
    b3 = row_vector(["x", "w"], [flat_vector(BIGINT, rng.integers(0, 2**50, 500)), flat_vector(BIGINT, np.arange(500))])
    p3 = row_vector(["a"], [flat_vector(BIGINT, np.concatenate([b3.columns[0].values[:200], rng.integers(0, 2**50, 800)]))])
    bb3 = PlanBuilder().values(b3.names, b3.types, source=1)
    (st,) = check_plan(PlanBuilder().values(p3.names, p3.types, source=0).hashJoin(["a"], ["x"], bb3, "", ["a", "w"]).planNode(), [p3, b3])
    assert stat(st, "b200.joinTableMode") == 1
    # empty build side
    empty = row_vector(["bk", "bv", "bs"], [flat_vector(BIGINT, []), flat_vector(INTEGER, []), flat_vector(VARCHAR, [])])
    for jt, outs in (("inner", ["pk", "bv"]), ("left", ["pk", "bv"]), ("anti", ["pk"])):
        check_plan(PlanBuilder().values(probe.names, probe.types, source=0)
                   .hashJoin(["pk"], ["bk"], PlanBuilder().values(empty.names, empty.types, source=1), "", outs, joinType=jt).planNode(), [probe, empty])


@pytest.mark.parametrize("keys", [[], ["c1"], ["c5", "c1"]])
def test_distinct_aggregates(keys):
    """DISTINCT aggregates (exec/DistinctAggregations.cpp): the adapter splits the node into a distinct grouping over
    (keys, x) and the plain aggregation over its rows. BIGINT, DOUBLE (NaN / -0 are one value each) and BOOLEAN inputs, NULLs ignored."""
    rv = table(n=4000, seed=21)
    for col in ("c0", "c3", "c4"):
        aggs = [f"count(distinct {col})"] if col == "c4" else [f"count(distinct {col})", f"sum(distinct {col})", f"min(distinct {col})", f"max(distinct {col})", f"avg(distinct {col})"]
        check_plan(PlanBuilder().values(rv.names, rv.types).singleAggregation(keys, aggs).planNode(), [rv])
        check_plan(PlanBuilder().values(rv.names, rv.types).filter("c1 <> 3").singleAggregation(keys, aggs).planNode(), [rv], batch_rows=900)
    rng = np.random.default_rng(3)
    d = row_vector(["k", "x"], [flat_vector(BIGINT, rng.integers(0, 5, 3000)), flat_vector(DOUBLE, rng.choice([0.0, -0.0, NAN, 1.5, 2.5, INF], 3000).tolist())])
    check_plan(PlanBuilder().values(d.names, d.types).singleAggregation(["k"], ["count(distinct x)", "max(distinct x)"]).planNode(), [d])


def test_distinct_aggregates_unsupported_shapes():
    rv = table(n=100, seed=1)
    for aggs in (["count(distinct c0)", "sum(c2)"], ["count(distinct c0)", "count(distinct c1)"]):
        with pytest.raises(Exception, match="DISTINCT"):
            check_plan(PlanBuilder().values(rv.names, rv.types).singleAggregation(["c1"], aggs).planNode(), [rv])


def _keyed_join_tables(seed=5, n=3000, m=400):
    rng = np.random.default_rng(seed)
    def maybe(v, p=0.05):
        return [None if rng.random() < p else x for x in v]
    probe = row_vector(["pa", "pb", "pd", "ps", "pv"], [
        flat_vector(BIGINT, maybe((rng.integers(0, 40, n) * (2**40)).tolist())),
        flat_vector(BIGINT, maybe((rng.integers(0, 30, n) * (2**35) - 2**50).tolist())),
        flat_vector(DOUBLE, maybe(rng.choice([0.0, -0.0, 1.5, NAN, -2.25, 1e300, 7.0], n).tolist())),
        dictionary_vector(VARCHAR, rng.integers(0, 7, n), ["ash", "birch", "cedar", "elm", "fir", None, "oak"]),
        flat_vector(BIGINT, np.arange(n))])
    build = row_vector(["ba", "bb", "bd", "bs", "bw"], [
        flat_vector(BIGINT, maybe((rng.integers(0, 40, m) * (2**40)).tolist())),
        flat_vector(BIGINT, maybe((rng.integers(0, 30, m) * (2**35) - 2**50).tolist())),
        flat_vector(DOUBLE, maybe(rng.choice([0.0, 1.5, NAN, -2.25, 3.0], m).tolist())),
        dictionary_vector(VARCHAR, rng.integers(0, 6, m), ["oak", "elm", None, "fir", "yew", "elm"]),  # a repeated entry: one id
        flat_vector(BIGINT, np.arange(m) * 10)])
    return probe, build


@pytest.mark.parametrize("keys", [(["pd"], ["bd"]), (["ps"], ["bs"]), (["pa", "pb"], ["ba", "bb"]), (["ps", "pd", "pa"], ["bs", "bd", "ba"])])
@pytest.mark.parametrize("join_type", ["inner", "left", "semi", "anti"])
def test_hash_join_keyed_mode(keys, join_type):
    """The reference's kHash mode for joins (exec/HashTable.cpp:1751-1838): DOUBLE keys (NaN = NaN, -0 = +0), VARCHAR
    keys (different dictionaries on the two sides) and two wide BIGINT keys that do not pack into one normalized word."""
    probe, build = _keyed_join_tables()
    b = PlanBuilder().values(build.names, build.types, source=1)
    outs = ["pv", "ps", "pd"] if join_type in ("semi", "anti") else ["pv", "ps", "pd", "bw", "bs", "bd"]
    plan = PlanBuilder().values(probe.names, probe.types, source=0).hashJoin(keys[0], keys[1], b, "", outs, joinType=join_type).planNode()
    (st,) = check_plan(plan, [probe, build])
    assert stat(st, "b200.joinTableMode") == 2
    check_plan(plan, [probe, build], batch_rows=700)  # multi-batch build side (dictionaries merged) and probe side


def test_hash_join_keyed_mode_filter_and_unique():
    probe, build = _keyed_join_tables(seed=9)
    b = PlanBuilder().values(build.names, build.types, source=1)
    plan = PlanBuilder().values(probe.names, probe.types, source=0).hashJoin(["pa", "pb"], ["ba", "bb"], b, "pv % 3 = 0 and bw > 100", ["pv", "bw"]).planNode()
    check_plan(plan, [probe, build])
    # unique DOUBLE build keys: the one-pass probe
    rng = np.random.default_rng(2)
    bu = row_vector(["bd", "bw"], [flat_vector(DOUBLE, (rng.permutation(500) * 0.25).tolist() + [NAN]), flat_vector(BIGINT, np.arange(501))])
    pu = row_vector(["pd"], [flat_vector(DOUBLE, (rng.integers(0, 800, 4000) * 0.25).tolist() + [NAN, None])])
    bb = PlanBuilder().values(bu.names, bu.types, source=1)
    (st,) = check_plan(PlanBuilder().values(pu.names, pu.types, source=0).hashJoin(["pd"], ["bd"], bb, "", ["pd", "bw"]).planNode(), [pu, bu])
    assert stat(st, "b200.uniqueKeyProbes") > 0 and stat(st, "b200.joinTableMode") == 2


@pytest.mark.parametrize("n", [50_000])
def test_tpch_q14_fused_and_generic(n):
    nparts = 3000
    h, col = _lineitem(n, seed=13, nparts=nparts)
    li = row_vector(["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"], [col(c) for c in ["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"]])
    part = {k: v.numpy() for k, v in tpch.gen_part(nparts, seed=5).items()}
    pt = row_vector(["p_partkey", "p_type"], [flat_vector(BIGINT, part["p_partkey"]), dictionary_vector(VARCHAR, part["p_type"], tpch.PTYPE_DICT)])
    st_f, st_g = check_plan(q14_plan(li, pt), [li, pt], configs=(FUSED, GENERIC), rel_tol=1e-12)
    assert stat(st_f, "b200.fusedBatches") == 1 and stat(st_g, "b200.fusedBatches") == 0
    check_plan(q14_plan(li, pt), [li, pt], configs=(FUSED, GENERIC), batch_rows=20_000, rel_tol=1e-12)


def test_upload_cache_shares_host_buffers_between_tasks():
    """vb2_upload_cache: a second task over the same host buffers reuses the resident device copies
    (task.h2dBytes counts only what was actually copied); results are unchanged."""
    from velox_b200.task import Task, UploadCache
    n = 200_000
    rng = np.random.default_rng(3)
    rv = row_vector(["k", "a", "b"], [flat_vector(INTEGER, rng.integers(0, 50, n).astype(np.int32)), flat_vector(DOUBLE, rng.standard_normal(n)),
                                      flat_vector(BIGINT, rng.integers(0, 1000, n))])
    p1 = PlanBuilder().values(rv.names, rv.types).filter("a < 0.5").singleAggregation(["k"], ["sum(b)", "count(0)"]).planNode()  # integer results: exact
    p2 = PlanBuilder().values(rv.names, rv.types).project(["a * 2.0 AS x", "b"]).singleAggregation([], ["sum(x)", "sum(b)"]).planNode()
    want1, want2 = check_plan(p1, [rv]), check_plan(p2, [rv])  # parity of both plans on their own
    del want1, want2
    cache = UploadCache()
    copied = []
    outs = []
    for plan in (p1, p2, p1):
        t = Task(plan)
        t.set_upload_cache(cache)
        t.add_input(0, rv)
        outs.append(t.run().rows())
        copied.append(t.stats()["task.h2dBytes"])
        t.close()
    cache.close()
    full = n * (4 + 8 + 8)
    assert copied[0] >= full and copied[1] < full // 10 and copied[2] < full // 10, copied
    assert sorted(outs[0]) == sorted(outs[2])
    # without a cache every task copies everything again
    t = Task(p1)
    t.add_input(0, rv)
    assert sorted(t.run().rows()) == sorted(outs[0])
    assert t.stats()["task.h2dBytes"] >= full
    t.close()


def test_exchange_operators_single_rank():
    """PartitionedOutput -> Exchange inside one task (world 1: every partition is local): the shuffle
    is the identity on the row multiset, whatever the encodings, NULLs and batch splits."""
    rv = table(n=4000, seed=21)
    plan = (PlanBuilder().values(rv.names, rv.types).filter("c1 < 8").project(["c0", "c1", "c2 * 2.0 AS d", "c5", "c4"])
            .partitionedOutput(["c0"]).planNode())
    check_plan(plan, [rv])
    check_plan(plan, [rv], batch_rows=900)
    plan = (PlanBuilder().values(rv.names, rv.types).partialAggregation(["c1", "c5"], ["sum(c2)", "avg(c3)", "count(0)", "max(c0)"])
            .gatherExchange().finalAggregation().planNode())
    check_plan(plan, [rv], rel_tol=1e-11)
    check_plan(plan, [rv], batch_rows=700, rel_tol=1e-11)
    plan = PlanBuilder().values(rv.names, rv.types).partitionedOutputBroadcast().singleAggregation([], ["count(0)", "sum(c0)"]).planNode()
    check_plan(plan, [rv])


def test_q14_multi_fragment_plan_single_rank():
    """The distributed Q14 plan (both join sides behind an exchange, gathered partial aggregates) on one rank."""
    import bench
    n, nparts = 60_000, 3000
    h, col = _lineitem(n, seed=14, nparts=nparts)
    names1 = ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate"]
    rv1 = row_vector(names1, [col(c) for c in names1])
    names14 = ["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"]
    li = row_vector(names14, [col(c) for c in names14])
    part = {k: v.numpy() for k, v in tpch.gen_part(nparts, seed=5).items()}
    pt = row_vector(["p_partkey", "p_type"], [flat_vector(BIGINT, part["p_partkey"]), dictionary_vector(VARCHAR, part["p_type"], tpch.PTYPE_DICT)])
    d1, d14 = bench.plans_distributed(rv1, li, pt)
    st_f, st_g = check_plan(d14, [li, pt], configs=(FUSED, GENERIC), rel_tol=1e-12)
    assert stat(st_f, "b200.fusedBatches") == 1, "the probe after the exchange must take the fused kernel"
    check_plan(d1, [rv1], configs=(FUSED, GENERIC), rel_tol=1e-12)
    check_plan(d1, [rv1], configs=(FUSED,), batch_rows=25_000, rel_tol=1e-12)


def test_group_keys_with_reordered_dictionaries_between_batches():
    """Two batches whose VARCHAR key dictionaries list the same strings in a different order (Arrow
    record batches do this): the per-dictionary id LUT must follow the dictionary, not its address."""
    from velox_b200.task import Task
    a = row_vector(["k", "v"], [dictionary_vector(VARCHAR, np.array([0, 1, 2, 0], dtype=np.int32), ["x", "y", "z"]), flat_vector(BIGINT, [1, 10, 100, 1000])])
    b = row_vector(["k", "v"], [dictionary_vector(VARCHAR, np.array([0, 1, 2, 0], dtype=np.int32), ["z", "x", "y"]), flat_vector(BIGINT, [2, 20, 200, 2000])])
    plan = PlanBuilder().values(a.names, a.types).singleAggregation(["k"], ["sum(v)", "count(0)"]).planNode()
    for _ in range(20):  # allocator reuse makes a stale pointer-keyed cache hit likely within a few rounds
        t = Task(plan)
        t.add_input(0, a)
        t.add_input(0, b)
        got = {r[0]: (r[1], r[2]) for r in t.run().rows()}
        t.close()
        assert got == {"x": (1 + 1000 + 20, 3), "y": (10 + 200, 2), "z": (100 + 2 + 2000, 3)}, got


def test_partial_final_collapse_keeps_results():
    """partial -> final in one driver is run as a single aggregation (adapter); with the fused
    pipelines off the pair stays apart. Both must give the oracle's answer."""
    rv = table(n=5000, seed=33)
    plan = (PlanBuilder().values(rv.names, rv.types).filter("c1 < 9").project(["c1", "c5", "c2", "c3 + 1.0 AS e", "c0"])
            .partialAggregation(["c1", "c5"], ["sum(c2)", "avg(e)", "count(0)", "min(c0)", "max(c2)"]).localPartition([]).finalAggregation().planNode())
    st_f, st_g = check_plan(plan, [rv], configs=(FUSED, GENERIC), rel_tol=1e-11)
    nagg = lambda st: len([k for k in st if k.endswith("B200HashAggregation.inputPositions")])
    assert nagg(st_f) == 1 and nagg(st_g) == 2


def test_pipeline_jit_shapes_without_an_aot_kernel():
    """Plan shapes with no ahead-of-time fused kernel are instantiated from the expression templates
    with NVRTC (fused_jit.cu) and still run as ONE scan -> filter -> project -> aggregate kernel:
    Q6 with a fourth predicate, and a grouped sum/avg over different expressions."""
    n = 300_000
    h, col = _lineitem(n, seed=77)
    names = ["l_shipdate", "l_extendedprice", "l_quantity", "l_discount", "l_tax", "l_linestatus"]
    rv = row_vector(names, [col(c) for c in names])
    tol = max(1e-12, n * 2.0 ** -53)
    q6x = (PlanBuilder().values(rv.names, rv.types)
           .filter("l_shipdate between '1994-01-01'::DATE and '1994-12-31'::DATE and l_discount between 0.05 and 0.07 and l_quantity < 24.0 "
                   "and l_extendedprice > 1000.0")
           .project(["l_extendedprice * l_discount AS r"]).singleAggregation([], ["sum(r)", "count(0)"]).planNode())
    st_f, st_g = check_plan(q6x, [rv], configs=(FUSED, GENERIC), rel_tol=tol, oracle_batch_rows=100_000)
    assert stat(st_f, "b200.fusedBatches") == 1 and stat(st_g, "b200.fusedBatches") == 0
    grouped = (PlanBuilder().values(rv.names, rv.types).filter("l_tax <= 0.06 and l_quantity >= 3.0")
               .project(["l_linestatus", "l_extendedprice / (1.0 + l_tax) AS net", "l_quantity - l_discount AS q"])
               .singleAggregation(["l_linestatus"], ["sum(net)", "avg(q)", "count(0)"]).planNode())
    st_f, st_g = check_plan(grouped, [rv], configs=(FUSED, GENERIC), rel_tol=tol, oracle_batch_rows=100_000)
    assert stat(st_f, "b200.fusedBatches") == 1 and stat(st_g, "b200.fusedBatches") == 0


def test_late_materialization_matches_full_scan():
    """Selective filter: the filter-first path (bitmap -> row numbers -> gather-aggregate) and the full
    fused scan give the same answer as the oracle; the planner picks it from a sampled selectivity."""
    n, nparts = 400_000, 5000
    h, col = _lineitem(n, seed=78, nparts=nparts)
    li = row_vector(["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"], [col(c) for c in ["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"]])
    part = {k: v.numpy() for k, v in tpch.gen_part(nparts, seed=5).items()}
    pt = row_vector(["p_partkey", "p_type"], [flat_vector(BIGINT, part["p_partkey"]), dictionary_vector(VARCHAR, part["p_type"], tpch.PTYPE_DICT)])
    late = {"b200.late_materialization_min_rows": "1000"}
    full = {"b200.late_materialization": "false"}
    st_l, st_f = check_plan(q14_plan(li, pt), [li, pt], configs=(late, full), rel_tol=1e-12, oracle_batch_rows=100_000)
    assert stat(st_l, "b200.selectiveBatches") == 1 and stat(st_f, "b200.selectiveBatches") == 0
    assert 5_000 < stat(st_l, "b200.sampledSelectivityPpm") < 25_000
    # several batches: the decision of the first batch holds for the rest
    (st,) = check_plan(q14_plan(li, pt), [li, pt], configs=(late,), batch_rows=100_000, rel_tol=1e-12, oracle_batch_rows=100_000)
    assert stat(st, "b200.selectiveBatches") == 4


def test_filter_project_fast_filter_kernel():
    """A FilterProject that is not absorbed into an aggregation evaluates a flat NULL-free filter with
    the TMA-staged bitmap kernel; rows and their order equal the oracle's."""
    n = 200_000
    h, col = _lineitem(n, seed=79)
    names = ["l_shipdate", "l_extendedprice", "l_discount", "l_partkey"]
    rv = row_vector(names, [col(c) for c in names])
    plan = (PlanBuilder().values(rv.names, rv.types).filter("l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE")
            .project(["l_extendedprice * (1.0 - l_discount) AS rev", "l_partkey"]).planNode())
    st_f, st_g = check_plan(plan, [rv], configs=(FUSED, GENERIC), oracle_batch_rows=100_000)
    assert stat(st_f, "b200.fastFilterBatches") == 1 and stat(st_g, "b200.fastFilterBatches") == 0


def test_group_by_double_keys():
    """DOUBLE grouping keys (keyed hash table, the reference's kHash mode): all NaNs are one group, -0.0
    and +0.0 are one group, NULL is a group (velox/type/FloatingPointUtil.h equality)."""
    rng = np.random.default_rng(91)
    n = 5000
    d = rng.integers(-20, 20, n).astype(float) / 4.0
    d[rng.random(n) < 0.05] = NAN
    d[::97] = -0.0
    d[1::97] = 0.0
    dn = rng.random(n) < 0.04
    rv = row_vector(["d", "k", "v", "x"], [flat_vector(DOUBLE, d, dn), flat_vector(INTEGER, rng.integers(0, 3, n).astype(np.int32)),
                                         flat_vector(BIGINT, rng.integers(-1000, 1000, n)), flat_vector(DOUBLE, rng.standard_normal(n))])
    plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(["d"], ["sum(v)", "count(0)", "avg(x)", "min(v)"]).planNode()
    (st,) = check_plan(plan, [rv], rel_tol=1e-11)
    assert stat(st, "b200.aggMode") == 3  # keyed
    check_plan(plan, [rv], batch_rows=700, rel_tol=1e-11)
    plan = (PlanBuilder().values(rv.names, rv.types).partialAggregation(["k", "d"], ["sum(v)", "max(x)", "count(0)"]).intermediateAggregation()
            .finalAggregation().planNode())
    check_plan(plan, [rv], configs=(GENERIC,), batch_rows=900, rel_tol=1e-11)


def test_group_by_keys_wider_than_one_normalized_word():
    """Two BIGINT keys spanning most of the int64 range do not fit one 64-bit normalized key: the groups
    move to a keyed table (kHash, exec/HashTable.cpp:1751-1838) — also when the wide values only show
    up in a later batch, after array-mode groups exist."""
    rng = np.random.default_rng(92)
    n = 6000
    a = rng.integers(0, 5, n)
    b = rng.integers(0, 7, n)
    a[3000:] = rng.choice(np.array([-(2 ** 62), 2 ** 62, 12345678901234, -7, 0]), n - 3000)   # wide values arrive late
    b[3000:] = rng.choice(np.array([2 ** 61, -(2 ** 61) - 5, 3, 99]), n - 3000)
    an = rng.random(n) < 0.03
    s = dictionary_vector(VARCHAR, rng.integers(0, 4, n).astype(np.int32), ["p", "q", "r", "s"])
    rv = row_vector(["a", "b", "s", "v"], [flat_vector(BIGINT, a, an), flat_vector(BIGINT, b), s, flat_vector(DOUBLE, rng.standard_normal(n))])
    plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(["a", "b", "s"], ["sum(v)", "count(0)", "max(v)"]).planNode()
    (st,) = check_plan(plan, [rv], batch_rows=1000, rel_tol=1e-11)
    assert stat(st, "b200.aggMode") == 3 and stat(st, "b200.aggRelayouts") >= 2
    check_plan(plan, [rv], rel_tol=1e-11)


def test_hash_build_many_batches_with_nulls_booleans_and_mixed_dictionaries():
    """HashBuild::addInput appends every batch (exec/HashBuild.cpp:442-598): build sides that arrive in
    many batches with NULLs, BOOLEAN columns and VARCHAR dictionaries that differ from batch to batch."""
    from velox_b200.task import Task
    from oracle import pyoracle
    from util import assert_equal_results
    rng = np.random.default_rng(93)
    probe = row_vector(["pk", "pv"], [flat_vector(BIGINT, rng.integers(0, 60, 4000), rng.random(4000) < 0.05), flat_vector(DOUBLE, rng.standard_normal(4000))])
    alphabets = [["red", "green", "blue"], ["blue", "black", "red", "white"], ["green"]]
    builds = []
    for i, alpha in enumerate(alphabets):
        m = 70
        builds.append(row_vector(["bk", "flag", "name", "bv"],
                                 [flat_vector(BIGINT, rng.integers(0, 60, m), rng.random(m) < 0.1), flat_vector(BOOLEAN, rng.random(m) < 0.5, rng.random(m) < 0.1),
                                  dictionary_vector(VARCHAR, rng.integers(0, len(alpha), m).astype(np.int32), alpha), flat_vector(INTEGER, rng.integers(0, 9, m).astype(np.int32), rng.random(m) < 0.2)]))
    for jt, outs in (("inner", ["pk", "pv", "flag", "name", "bv"]), ("left", ["pk", "flag", "name", "bv"])):
        plan = (PlanBuilder().values(probe.names, probe.types, source=0)
                .hashJoin(["pk"], ["bk"], PlanBuilder().values(builds[0].names, builds[0].types, source=1), "", outs, joinType=jt).planNode())
        t = Task(plan)
        t.add_input(0, probe)
        for b in builds:
            t.add_input(1, b)
        got = t.run()
        t.close()
        # the oracle gets the same build side as one table (strings flattened)
        cat = lambda i, ty: flat_vector(ty, [v for b in builds for v in b.columns[i].to_pylist()])
        whole = row_vector(builds[0].names, [cat(0, BIGINT), cat(1, BOOLEAN), cat(2, VARCHAR), cat(3, INTEGER)])
        want = pyoracle.run_plan(plan, [probe, whole])
        assert_equal_results(got, want)


def test_aggregation_radix_partitioned_batches():
    """High-cardinality batches are radix-partitioned by the top bits of their table hash before the
    find-or-insert pass (radix_partition.cu): one or two keys, several aggregates over 4- and 8-byte
    inputs, several batches into the same table; the table is sized from the HyperLogLog estimate."""
    n = 500_000
    rng = np.random.default_rng(12)
    pool_keys = rng.integers(0, 2**40, 90_000)
    keys = pool_keys[rng.integers(0, 90_000, n)]
    rv = row_vector(["k", "v", "w", "x"], [flat_vector(BIGINT, keys), flat_vector(BIGINT, (np.arange(n) % 1000)), flat_vector(INTEGER, rng.integers(-50, 50, n).astype(np.int32)),
                                         flat_vector(DOUBLE, rng.standard_normal(n))])
    part = {"b200.agg_radix_partition": "true", "b200.agg_partition_min_rows": "1000"}
    plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(["k"], ["sum(v)", "count(0)", "min(w)", "max(x)", "avg(x)"]).planNode()
    (st,) = check_plan(plan, [rv], configs=(part,), rel_tol=1e-11, oracle_batch_rows=100_000)
    assert stat(st, "b200.partitionedBatches") == 1 and stat(st, "b200.aggMode") == 2
    (st,) = check_plan(plan, [rv], configs=(part,), batch_rows=125_000, rel_tol=1e-11, oracle_batch_rows=100_000)
    assert stat(st, "b200.partitionedBatches") == 4
    two = row_vector(["a", "b", "v"], [flat_vector(BIGINT, keys), flat_vector(INTEGER, (keys % 7).astype(np.int32), rng.random(n) < 0.01), flat_vector(DOUBLE, np.ones(n))])
    plan2 = PlanBuilder().values(two.names, two.types).singleAggregation(["a", "b"], ["sum(v)", "count(0)"]).planNode()
    (st,) = check_plan(plan2, [two], configs=(part,), oracle_batch_rows=100_000)
    assert stat(st, "b200.partitionedBatches") == 1
    # partial -> final: the final step merges intermediate columns through the same path
    plan3 = (PlanBuilder().values(rv.names, rv.types).partialAggregation(["k"], ["sum(v)", "avg(x)", "count(0)"]).intermediateAggregation()
             .finalAggregation().planNode())
    check_plan(plan3, [rv], configs=(dict(part, **GENERIC),), batch_rows=250_000, rel_tol=1e-11, oracle_batch_rows=100_000)
