"""Anchors the oracle (and through it the GPU path) on the reference's own data generator and the
published TPC-H answers (CPU only):
  * oracle/_ref/libtpchref.so = the reference's vendored dbgen compiled from /root/reference; its
    rows reproduce the goldens of velox/tpch/gen/tests/TpchGenTest.cpp:326-369;
  * the oracle's Q1 / Q6 / Q14 over that SF1 data equal the TPC-H qualification answers
    (tests/golden/tpch_sf1_answers.json) — counts exactly, sums to the published digits;
  * the committed SF0.01 fixture (used by the GPU tests) is exactly what the generator produces."""
import datetime
import json
import os

import numpy as np
import pytest

from oracle import pyoracle, tpch_ref
from velox_b200.plan import PlanBuilder
from velox_b200.vector import BIGINT, DOUBLE, INTEGER, VARCHAR, dictionary_vector, flat_vector, row_vector

HERE = os.path.dirname(os.path.abspath(__file__))
ANS = json.load(open(os.path.join(HERE, "golden", "tpch_sf1_answers.json")))
needs_ref = pytest.mark.skipif(not tpch_ref.available(), reason="oracle/_ref not built (needs /root/reference at build time)")


def iso(days):
    return (datetime.date(1970, 1, 1) + datetime.timedelta(days=int(days))).isoformat()


def lineitem_vectors(li, names):
    cols = []
    for n in names:
        if n in ("l_returnflag", "l_linestatus"):
            values, codes = np.unique(li[n], return_inverse=True)
            cols.append(dictionary_vector(VARCHAR, codes.astype(np.int32), [chr(v) for v in values]))
        elif n == "l_shipdate":
            cols.append(flat_vector(INTEGER, li[n]))
        elif n in ("l_partkey", "l_orderkey"):
            cols.append(flat_vector(BIGINT, li[n]))
        else:
            cols.append(flat_vector(DOUBLE, li[n]))
    return row_vector(names, cols)


def tpch_plans(rv1, rv6, rv14, pt):
    q1 = (PlanBuilder().values(rv1.names, rv1.types).filter("l_shipdate < '1998-09-03'::DATE")
          .project(["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_extendedprice * (1.0 - l_discount) AS l_sum_disc_price",
                    "l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS l_sum_charge", "l_discount"])
          .partialAggregation(["l_returnflag", "l_linestatus"],
                              ["sum(l_quantity)", "sum(l_extendedprice)", "sum(l_sum_disc_price)", "sum(l_sum_charge)", "avg(l_quantity)",
                               "avg(l_extendedprice)", "avg(l_discount)", "count(0)"]).localPartition([]).finalAggregation().planNode())
    q6 = (PlanBuilder().values(rv6.names, rv6.types)
          .filter("l_shipdate between '1994-01-01'::DATE and '1994-12-31'::DATE and l_discount between 0.05 and 0.07 and l_quantity < 24.0")
          .project(["l_extendedprice * l_discount"]).partialAggregation([], ["sum(p0)"]).localPartition([]).finalAggregation().planNode())
    build = PlanBuilder().values(pt.names, pt.types, source=1)
    q14 = (PlanBuilder().values(rv14.names, rv14.types, source=0).filter("l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE")
           .project(["l_extendedprice * (1.0 - l_discount) as part_revenue", "l_shipdate", "l_partkey"])
           .hashJoin(["l_partkey"], ["p_partkey"], build, "", ["part_revenue", "p_type"])
           .project(["(CASE WHEN (p_type LIKE 'PROMO%') THEN part_revenue ELSE 0.0 END) as filter_revenue", "part_revenue"])
           .partialAggregation([], ["sum(part_revenue) as total_revenue", "sum(filter_revenue) as total_promo_revenue"])
           .localPartition([]).finalAggregation().project(["100.00 * total_promo_revenue/total_revenue as promo_revenue"]).planNode())
    return q1, q6, q14


@needs_ref
def test_dbgen_reproduces_tpchgen_goldens():
    g = ANS["tpchgen_goldens"]
    li = tpch_ref.gen_lineitem(1.0, *g["batch1"]["orders"])
    n = len(li["l_orderkey"])
    assert 100 <= n <= 700
    assert [int(li["l_orderkey"][0]), float(li["l_quantity"][0]), iso(li["l_shipdate"][0])] == g["batch1"]["first"]
    assert [int(li["l_orderkey"][-1]), float(li["l_quantity"][-1]), iso(li["l_shipdate"][-1])] == g["batch1"]["last"]
    li = tpch_ref.gen_lineitem(1.0, *g["batch2"]["orders"])
    assert [int(li["l_orderkey"][0]), iso(li["l_shipdate"][0])] == g["batch2"]["first"]
    assert [int(li["l_orderkey"][-1]), iso(li["l_shipdate"][-1])] == g["batch2"]["last"]


@needs_ref
def test_oracle_matches_published_tpch_answers_sf1():
    li = tpch_ref.gen_lineitem(1.0)
    part = tpch_ref.gen_part(1.0)
    assert len(li["l_orderkey"]) == ANS["lineitem_rows"]
    rv1 = lineitem_vectors(li, ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate"])
    rv6 = lineitem_vectors(li, ["l_shipdate", "l_extendedprice", "l_quantity", "l_discount"])
    rv14 = lineitem_vectors(li, ["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"])
    types, codes = np.unique(np.array(part["p_type"]), return_inverse=True)
    pt = row_vector(["p_partkey", "p_type"], [flat_vector(BIGINT, part["p_partkey"]), dictionary_vector(VARCHAR, codes.astype(np.int32), types.tolist())])
    q1, q6, q14 = tpch_plans(rv1, rv6, rv14, pt)
    threads = min(8, os.cpu_count() or 1)
    got1 = {r[0] + r[1]: r[2:] for r in pyoracle.run_plan(q1, [rv1], threads=threads).rows()}
    assert set(got1) == {"AF", "NF", "NO", "RF"}
    for key, want in ANS["q1"].items():
        if key == "columns":
            continue
        g = got1[key]
        assert g[7] == want[7]                                  # count: exact
        for a, b in zip(g[:7], want[:7]):
            assert a == pytest.approx(b, rel=1e-11)              # published digits
    assert pyoracle.run_plan(q6, [rv6], threads=threads).rows()[0][0] == pytest.approx(ANS["q6"], rel=1e-12)
    assert pyoracle.run_plan(q14, [rv14, pt], threads=threads).rows()[0][0] == pytest.approx(ANS["q14"], rel=1e-12)
    # single driver = the reference's exact sequential accumulation order
    assert pyoracle.run_plan(q6, [rv6], threads=1).rows()[0][0] == pytest.approx(ANS["q6"], rel=1e-12)


@needs_ref
def test_committed_fixture_is_what_the_generator_produces():
    fx = np.load(os.path.join(HERE, "golden", "tpch_sf001.npz"))
    li = tpch_ref.gen_lineitem(0.01)
    for k, v in li.items():
        assert np.array_equal(fx[k], v), k
    pt = tpch_ref.gen_part(0.01)
    assert np.array_equal(fx["p_partkey"], pt["p_partkey"])
    assert [str(fx["p_type_dict"][c]) for c in fx["p_type_codes"]] == pt["p_type"]
