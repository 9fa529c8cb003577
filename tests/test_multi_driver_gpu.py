"""Several drivers per pipeline (task.max_drivers = Task::start's maxDrivers): sibling B200 operators run
concurrently on their own threads and streams, share state only through Task::allPeersFinished
(velox/exec/Task.cpp:2451, used by HashBuild::finishHashBuild exec/HashBuild.cpp:819) and JoinBridges,
and meet again at a LocalPartition gather (exec/LocalPartition.cpp). Results must equal the oracle's and
the single-driver run's. Modelled on the reference's multi-threaded variants of AggregationTest /
HashJoinTest (AssertQueryBuilder::maxDrivers)."""
import numpy as np
import pytest

from util import assert_equal_results, stat
from oracle import pyoracle
from velox_b200.plan import PlanBuilder
from velox_b200.task import run_plan
from velox_b200.vector import BIGINT, DOUBLE, INTEGER, VARCHAR, dictionary_vector, flat_vector, row_vector

pytestmark = pytest.mark.gpu
GENERIC = {"b200.fused_pipelines": "false"}


def lineitem(n, seed=0):
    rng = np.random.default_rng(seed)
    return row_vector(["k", "flag", "qty", "price", "disc", "ship"], [
        flat_vector(BIGINT, rng.integers(1, 2001, n)),
        dictionary_vector(VARCHAR, rng.integers(0, 3, n), ["A", "N", "R"]),
        flat_vector(DOUBLE, rng.integers(1, 51, n).astype(np.float64)),
        flat_vector(DOUBLE, np.round(rng.uniform(900, 2100, n), 2)),
        flat_vector(DOUBLE, rng.integers(0, 11, n) / 100.0),
        flat_vector(INTEGER, rng.integers(8000, 10500, n).astype(np.int32))])


@pytest.mark.parametrize("drivers", [2, 4])
@pytest.mark.parametrize("cfg", [{}, GENERIC])
def test_partial_aggregation_on_n_drivers(drivers, cfg):
    rv = lineitem(200_000)
    plan = (PlanBuilder().values(rv.names, rv.types).filter("ship < 10000").project(["flag", "qty", "price * (1.0 - disc) as rev", "k"])
            .partialAggregation(["flag"], ["sum(qty)", "sum(rev)", "avg(rev)", "count(0)", "max(k)"]).localPartition([]).finalAggregation().orderBy(["flag"]).planNode())
    want = pyoracle.run_plan(plan, [rv], threads=1, batch_rows=10000)
    got, st = run_plan(plan, [rv], config=dict(cfg, **{"task.max_drivers": str(drivers)}), batch_rows=20_000)
    assert_equal_results(got, want, rel_tol=1e-11)
    assert st["task.numDrivers"] == drivers + 1                      # N partial drivers + the gather's consumer
    assert stat(st, "B200HashAggregation.inputPositions") >= 100_000  # the partial siblings' inputs add up under one key
    one, _ = run_plan(plan, [rv], config=cfg, batch_rows=20_000)
    assert_equal_results(got, one, rel_tol=1e-11)


@pytest.mark.parametrize("drivers", [2, 3])
def test_join_build_and_probe_on_n_drivers(drivers):
    """N build drivers (peers meet in Task::allPeersFinished, the last one builds the table from all rows),
    N probe drivers sharing the bridge's table, then the gather."""
    li = lineitem(120_000, seed=1)
    rng = np.random.default_rng(2)
    part = row_vector(["p", "ptype"], [flat_vector(BIGINT, np.arange(1, 1501)), dictionary_vector(VARCHAR, rng.integers(0, 4, 1500), ["PROMO X", "STD", "PROMO Y", "ECO"])])
    build = PlanBuilder().values(part.names, part.types, source=1)
    plan = (PlanBuilder().values(li.names, li.types, source=0).filter("ship between 8500 and 9500").project(["price * (1.0 - disc) as rev", "k"])
            .hashJoin(["k"], ["p"], build, "", ["rev", "ptype"])
            .project(["(CASE WHEN (ptype LIKE 'PROMO%') THEN rev ELSE 0.0 END) as promo", "rev"])
            .partialAggregation([], ["sum(rev) as t", "sum(promo) as p", "count(0) as c"]).localPartition([]).finalAggregation().planNode())
    want = pyoracle.run_plan(plan, [li, part], threads=1, batch_rows=10000)
    for cfg in ({}, GENERIC):
        got, st = run_plan(plan, [li, part], config=dict(cfg, **{"task.max_drivers": str(drivers)}), batch_rows=10_000)
        # `part` is one batch of 1500 rows and li twelve: split the build side too
        assert_equal_results(got, want, rel_tol=1e-11)
    from velox_b200.task import Task, split_rowvector
    t = Task(plan, {"task.max_drivers": str(drivers), "b200.fused_pipelines": "false"})
    try:
        for b in split_rowvector(li, 10_000):
            t.add_input(0, b)
        for b in split_rowvector(part, 250):
            t.add_input(1, b)
        got = t.run()
        st = t.stats()
    finally:
        t.close()
    assert_equal_results(got, want, rel_tol=1e-11)
    assert stat(st, "b200.buildPeers") == drivers


def test_high_cardinality_partial_final_on_4_drivers():
    n = 400_000
    rng = np.random.default_rng(7)
    rv = row_vector(["k", "v"], [flat_vector(BIGINT, rng.integers(0, 150_000, n) * 7919 - 10**9), flat_vector(BIGINT, rng.integers(0, 1000, n))])
    plan = PlanBuilder().values(rv.names, rv.types).partialAggregation(["k"], ["sum(v)", "count(0)"]).localPartition([]).finalAggregation().planNode()
    want = pyoracle.run_plan(plan, [rv], threads=1, batch_rows=10000)
    got, st = run_plan(plan, [rv], config={"task.max_drivers": "4"}, batch_rows=50_000)
    assert_equal_results(got, want)
    assert st["task.numDrivers"] == 5


def test_error_in_one_driver_fails_the_task():
    from velox_b200._lib import VeloxUserError
    big = 2**62
    rv = row_vector(["a"], [flat_vector(BIGINT, [1, 2, 3, big] * 1000)])
    plan = PlanBuilder().values(rv.names, rv.types).project(["a + a as b"]).partialAggregation([], ["count(0)"]).localPartition([]).finalAggregation().planNode()
    with pytest.raises(VeloxUserError):
        run_plan(plan, [rv], config={"task.max_drivers": "3"}, batch_rows=500)
