"""Independent cross-check of the CPU oracle against pyarrow / Acero (the reference
pins operator results through DuckDB at test time, which is absent here; pyarrow is the columnar
engine that is). Not Velox, so only SQL-level semantics are compared: filters, projections,
grouped aggregates with NULL handling, and inner / left / semi / anti hash joins. Runs on the CPU."""
import math

import numpy as np
import pyarrow as pa
import pyarrow.compute as pc
import pytest

from oracle import pyoracle
from velox_b200.arrow import row_vector_from_arrow, row_vector_to_arrow
from velox_b200.plan import PlanBuilder


def table(n, seed, null_p=0.1):
    rng = np.random.default_rng(seed)

    def with_nulls(values, typ):
        mask = rng.random(n) < null_p
        return pa.array(values, type=typ, mask=mask)

    return pa.table({
        "k1": with_nulls(rng.integers(0, 7, n), pa.int32()),
        "k2": with_nulls(rng.integers(-3, 4, n), pa.int64()),
        "s": pa.array(rng.choice(["apple", "banana", "cherry", None], n).tolist(), type=pa.string()).dictionary_encode(),
        "x": with_nulls(np.round(rng.normal(0, 100, n), 3), pa.float64()),
        "q": with_nulls(rng.integers(0, 50, n).astype(np.float64), pa.float64()),
        "v": with_nulls(rng.integers(-1000, 1000, n), pa.int64()),
    })


def norm(rows):
    def key(r):
        return tuple((0, 0) if v is None else (3, 0) if isinstance(v, float) and math.isnan(v) else (1, round(v, 6)) if isinstance(v, float)
                     else (1, v) if not isinstance(v, str) else (2, v) for v in r)
    return sorted(rows, key=key)


def assert_same(got, want, rel=1e-9):
    g, w = norm(got), norm(want)
    assert len(g) == len(w), (len(g), len(w))
    for a, b in zip(g, w):
        for x, y in zip(a, b):
            if x is None or y is None:
                assert x is None and y is None, (a, b)
            elif isinstance(y, float):
                assert (math.isnan(x) and math.isnan(y)) or math.isclose(x, y, rel_tol=rel, abs_tol=1e-9), (a, b)
            else:
                assert x == y, (a, b)


def arrow_rows(t):
    cols = [t.column(i).to_pylist() for i in range(t.num_columns)]
    return list(zip(*cols)) if cols else []


def test_arrow_round_trip():
    t = table(1000, 1)
    back = row_vector_to_arrow(row_vector_from_arrow(t))
    assert arrow_rows(back) == arrow_rows(t.cast(back.schema) if back.schema != t.schema else t) or arrow_rows(back) == arrow_rows(t)


@pytest.mark.parametrize("seed", [2, 3])
def test_filter_project_against_arrow(seed):
    t = table(5000, seed)
    rv = row_vector_from_arrow(t)
    plan = (PlanBuilder().values(rv.names, rv.types).filter("q < 24.0 AND (x > 0.0 OR v between -100 and 100)")
            .project(["x * (1.0 - q) AS a", "v + k2 AS b", "k1"]).planNode())
    got = pyoracle.run_plan(plan, [rv], threads=1, batch_rows=997).rows()
    keep = pc.and_kleene(pc.less(t["q"], 24.0), pc.or_kleene(pc.greater(t["x"], 0.0), pc.and_kleene(pc.greater_equal(t["v"], -100), pc.less_equal(t["v"], 100))))
    f = t.filter(keep)  # drops rows whose predicate is NULL or false
    want = pa.table({"a": pc.multiply(f["x"], pc.subtract(1.0, f["q"])), "b": pc.add(f["v"], f["k2"]), "k1": f["k1"]})
    assert_same(got, arrow_rows(want))


@pytest.mark.parametrize("keys", [[], ["k1"], ["k1", "k2"], ["s"], ["s", "k1"]])
def test_group_by_against_arrow(keys):
    t = table(20_000, 7)
    rv = row_vector_from_arrow(t)
    plan = (PlanBuilder().values(rv.names, rv.types)
            .singleAggregation(keys, ["sum(x)", "sum(v)", "count(x)", "count(0)", "min(v)", "max(x)", "avg(q)"]).planNode())
    got = pyoracle.run_plan(plan, [rv], threads=1, batch_rows=4096).rows()
    aggs = [("x", "sum"), ("v", "sum"), ("x", "count"), ([], "count_all"), ("v", "min"), ("x", "max"), ("q", "mean")]
    want = t.group_by(keys, use_threads=False).aggregate(aggs)
    # pyarrow puts the aggregates first, then the keys
    order = keys + [c for c in want.schema.names if c not in keys]
    assert_same(got, arrow_rows(want.select(order)))


@pytest.mark.parametrize("keys", [["q"], ["w1", "w2"], ["q", "s", "w1"]])
def test_group_by_keyed_shapes_against_arrow(keys):
    """Grouping keys that take the keyed (kHash) group table on the device: DOUBLE keys (NaN is one group, NULL another)
    and two wide BIGINT keys whose ranges do not pack into one normalized word."""
    t = table(20_000, 31)
    rng = np.random.default_rng(5)
    n = t.num_rows
    q = np.where(rng.random(n) < 0.05, np.nan, rng.integers(0, 12, n).astype(np.float64) * 0.25)
    t = t.set_column(t.schema.get_field_index("q"), "q", pa.array(q, type=pa.float64(), mask=rng.random(n) < 0.05))
    t = t.append_column("w1", pa.array(rng.integers(0, 9, n) * (2**41), type=pa.int64(), mask=rng.random(n) < 0.05))
    t = t.append_column("w2", pa.array(rng.integers(0, 7, n) * (2**38) - 2**52, type=pa.int64()))
    rv = row_vector_from_arrow(t)
    plan = PlanBuilder().values(rv.names, rv.types).singleAggregation(keys, ["sum(v)", "count(x)", "count(0)", "min(v)", "max(x)"]).planNode()
    got = pyoracle.run_plan(plan, [rv], threads=4, batch_rows=4096).rows()
    want = t.group_by(keys, use_threads=False).aggregate([("v", "sum"), ("x", "count"), ([], "count_all"), ("v", "min"), ("x", "max")])
    order = keys + [c for c in want.schema.names if c not in keys]
    assert_same(got, arrow_rows(want.select(order)))


@pytest.mark.parametrize("keys", [[], ["k1"], ["s", "k2"]])
def test_distinct_aggregates_against_arrow(keys):
    """count(DISTINCT x) / sum / min / max over the distinct values of a group (exec/DistinctAggregations.cpp) against
    pyarrow's count_distinct and a distinct-then-aggregate restatement in pyarrow."""
    t = table(20_000, 11)
    rv = row_vector_from_arrow(t)
    plan = (PlanBuilder().values(rv.names, rv.types)
            .singleAggregation(keys, ["count(distinct v)", "sum(distinct v)", "min(distinct v)", "max(distinct v)", "avg(distinct v)"]).planNode())
    got = pyoracle.run_plan(plan, [rv], threads=4, batch_rows=4096).rows()
    # distinct (keys, v) rows first, then plain aggregates over them; count_distinct straight from pyarrow as a second opinion
    d = t.select(keys + ["v"]).group_by(keys + ["v"], use_threads=False).aggregate([])
    want = d.group_by(keys, use_threads=False).aggregate([("v", "count"), ("v", "sum"), ("v", "min"), ("v", "max"), ("v", "mean")])
    order = keys + [c for c in want.schema.names if c not in keys]
    assert_same(got, arrow_rows(want.select(order)))
    direct = t.group_by(keys, use_threads=False).aggregate([("v", "count_distinct")])
    by_key = {tuple(r[:-1]): r[-1] for r in arrow_rows(direct.select(keys + ["v_count_distinct"]))}
    for r in got:
        assert by_key[tuple(r[:len(keys)])] == r[len(keys)]


@pytest.mark.parametrize("join_type,arrow_type", [("inner", "inner"), ("left", "left outer"), ("semi", "left semi"), ("anti", "left anti")])
def test_hash_join_against_arrow(join_type, arrow_type):
    rng = np.random.default_rng(11)
    n, m = 4000, 300
    probe = pa.table({"pk": pa.array(rng.integers(0, 400, n), type=pa.int64(), mask=rng.random(n) < 0.05), "pv": pa.array(rng.standard_normal(n))})
    build = pa.table({"bk": pa.array(rng.integers(0, 400, m), type=pa.int64(), mask=rng.random(m) < 0.05), "bv": pa.array(rng.integers(0, 9, m), type=pa.int64())})
    prv, brv = row_vector_from_arrow(probe), row_vector_from_arrow(build)
    out_cols = ["pk", "pv"] if join_type in ("semi", "anti") else ["pk", "pv", "bv"]
    plan = (PlanBuilder().values(prv.names, prv.types, source=0)
            .hashJoin(["pk"], ["bk"], PlanBuilder().values(brv.names, brv.types, source=1), "", out_cols, joinType=join_type).planNode())
    got = pyoracle.run_plan(plan, [prv, brv], threads=1, batch_rows=512).rows()
    j = probe.join(build, keys="pk", right_keys="bk", join_type=arrow_type, use_threads=False)
    want = j.select(out_cols)
    assert_same(got, arrow_rows(want))


@pytest.mark.parametrize("join_type,arrow_type", [("inner", "inner"), ("left", "left outer"), ("semi", "left semi"), ("anti", "left anti")])
@pytest.mark.parametrize("shape", ["wide_pair", "string", "string_and_int", "double"])
def test_wide_and_string_key_joins_against_arrow(join_type, arrow_type, shape):
    """The key shapes that take the keyed (kHash) join table on the device: two wide BIGINT keys that do not pack into
    one normalized word, VARCHAR keys with different dictionaries on the two sides, and a mix. NULL keys never match."""
    rng = np.random.default_rng(23)
    n, m = 5000, 400
    words = ["ash", "birch", "cedar", "elm", "fir", "oak", "yew", None]
    probe = pa.table({
        "pa": pa.array(rng.integers(0, 30, n) * (2**40), type=pa.int64(), mask=rng.random(n) < 0.05),
        "pb": pa.array(rng.integers(0, 20, n) * (2**35) - 2**50, type=pa.int64(), mask=rng.random(n) < 0.05),
        "ps": pa.array(rng.choice(words, n).tolist(), type=pa.string()).dictionary_encode(),
        # NaN keys match NaN in both engines; -0.0 is left out: the reference (and the oracle) equate it with +0.0, Arrow compares bits
        "pd": pa.array(rng.choice([0.0, 1.5, float("nan"), -2.25, 1e300, 7.0], n), type=pa.float64(), mask=rng.random(n) < 0.05),
        "pv": pa.array(np.arange(n), type=pa.int64())})
    build = pa.table({
        "ba": pa.array(rng.integers(0, 30, m) * (2**40), type=pa.int64(), mask=rng.random(m) < 0.05),
        "bb": pa.array(rng.integers(0, 20, m) * (2**35) - 2**50, type=pa.int64(), mask=rng.random(m) < 0.05),
        "bs": pa.array(rng.choice(words[::-1], m).tolist(), type=pa.string()).dictionary_encode(),
        "bd": pa.array(rng.choice([0.0, 1.5, float("nan"), -2.25, 3.0], m), type=pa.float64(), mask=rng.random(m) < 0.05),
        "bw": pa.array(np.arange(m) * 10, type=pa.int64())})
    pk, bk = {"wide_pair": (["pa", "pb"], ["ba", "bb"]), "string": (["ps"], ["bs"]), "string_and_int": (["ps", "pa"], ["bs", "ba"]), "double": (["pd"], ["bd"])}[shape]
    prv, brv = row_vector_from_arrow(probe), row_vector_from_arrow(build)
    out_cols = ["pv"] if join_type in ("semi", "anti") else ["pv", "bw"]
    plan = (PlanBuilder().values(prv.names, prv.types, source=0)
            .hashJoin(pk, bk, PlanBuilder().values(brv.names, brv.types, source=1), "", out_cols, joinType=join_type).planNode())
    got = pyoracle.run_plan(plan, [prv, brv], threads=4, batch_rows=700).rows()
    plain = lambda t: pa.table({c: (t[c].cast(pa.string()) if pa.types.is_dictionary(t[c].type) else t[c]) for c in t.column_names})
    j = plain(probe).join(plain(build), keys=pk, right_keys=bk, join_type=arrow_type, use_threads=False)
    assert_same(got, arrow_rows(j.select(out_cols)))


def test_case_like_cast_against_arrow():
    t = table(4000, 21)
    rv = row_vector_from_arrow(t)
    plan = (PlanBuilder().values(rv.names, rv.types)
            .project(["CASE WHEN s LIKE '%an%' THEN x ELSE 0.0 END AS a",
                      "CASE WHEN k1 > 3 THEN v END AS b",             # no ELSE: NULL
                      "CAST(k1 AS DOUBLE) * 0.5 AS c",
                      "CAST(q AS BIGINT) + k2 AS d",
                      "s LIKE 'b_n%' AS e",
                      "NOT (x < 0.0) OR k1 IS NULL AS f"]).planNode())
    got = pyoracle.run_plan(plan, [rv], threads=1, batch_rows=777).rows()
    s = t["s"].cast(pa.string())
    like_an = pc.match_like(s, "%an%")
    want = pa.table({
        # CASE treats a NULL condition as not taken
        "a": pc.if_else(pc.fill_null(like_an, False), t["x"], 0.0),
        "b": pc.if_else(pc.fill_null(pc.greater(t["k1"], 3), False), t["v"], pa.scalar(None, pa.int64())),
        "c": pc.multiply(pc.cast(t["k1"], pa.float64()), 0.5),
        "d": pc.add(pc.cast(pc.round(t["q"]), pa.int64()), t["k2"]),
        "e": pc.match_like(s, "b_n%"),
        "f": pc.or_kleene(pc.invert(pc.less(t["x"], 0.0)), pc.is_null(t["k1"])),
    })
    assert_same(got, arrow_rows(want))


def test_order_by_against_arrow():
    """OrderBy exists in the oracle ahead of the device operator: directions,
    NULLS FIRST / LAST, NaN-largest doubles, strings, multi-key — against pyarrow's sort."""
    rng = np.random.default_rng(5)
    n = 3000
    x = np.round(rng.normal(0, 10, n), 1)
    x[rng.random(n) < 0.05] = np.nan
    t = pa.table({
        "id": pa.array(np.arange(n), type=pa.int64()),  # unique tiebreaker: the order is total
        "k": pa.array(rng.integers(0, 20, n), type=pa.int32(), mask=rng.random(n) < 0.1),
        "x": pa.array(x, type=pa.float64(), mask=rng.random(n) < 0.1),
        "s": pa.array(rng.choice(["a", "ab", "b", "", "zz"], n).tolist(), type=pa.string()),
    })
    rv = row_vector_from_arrow(t)
    cases = [
        (["k", "id"], [("k", "ascending"), ("id", "ascending")], "at_end"),
        (["k DESC NULLS FIRST", "id DESC"], [("k", "descending"), ("id", "descending")], "at_start"),
        (["s DESC", "id"], [("s", "descending"), ("id", "ascending")], "at_end"),
    ]
    for keys, sort_keys, null_placement in cases:
        plan = PlanBuilder().values(rv.names, rv.types).orderBy(keys).planNode()
        got = pyoracle.run_plan(plan, [rv], threads=2, batch_rows=700).rows()
        want = arrow_rows(t.take(pc.sort_indices(t, sort_keys=sort_keys, null_placement=null_placement)))
        assert [r[0] for r in got] == [r[0] for r in want], keys
    # doubles: NaN sorts as the largest value, NULLs last
    plan = PlanBuilder().values(rv.names, rv.types).orderBy(["x", "id"]).planNode()
    got = pyoracle.run_plan(plan, [rv], threads=1, batch_rows=1000).rows()
    xs = [r[2] for r in got]
    nn = [v for v in xs if v is not None]
    assert all(v is None for v in xs[len(nn):])
    finite = [v for v in nn if not math.isnan(v)]
    assert all(math.isnan(v) for v in nn[len(finite):]) and finite == sorted(finite)
