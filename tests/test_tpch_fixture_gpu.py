"""TPC-H Q1 / Q6 / Q14 over real dbgen rows (tests/golden/tpch_sf001.npz, produced by the
reference's own generator — tests/golden/make_tpch_fixture.py) through the operator-level C ABI,
fused and generic paths, against the oracle that tests/test_tpch_reference_data.py pins to the
published TPC-H answers. The part fixture holds 2 000 of the 200 000 part keys lineitem refers to,
so the join also exercises misses."""
import os

import numpy as np
import pytest

from test_tpch_reference_data import lineitem_vectors, tpch_plans
from util import FUSED, GENERIC, check_plan, stat
from velox_b200.vector import BIGINT, VARCHAR, dictionary_vector, flat_vector, row_vector

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def test_q1_q6_q14_on_dbgen_rows():
    fx = np.load(os.path.join(HERE, "golden", "tpch_sf001.npz"))
    li = {k: fx[k] for k in fx.files if k.startswith("l_")}
    rv1 = lineitem_vectors(li, ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate"])
    rv6 = lineitem_vectors(li, ["l_shipdate", "l_extendedprice", "l_quantity", "l_discount"])
    rv14 = lineitem_vectors(li, ["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"])
    pt = row_vector(["p_partkey", "p_type"], [flat_vector(BIGINT, fx["p_partkey"]),
                                             dictionary_vector(VARCHAR, fx["p_type_codes"], [str(s) for s in fx["p_type_dict"]])])
    q1, q6, q14 = tpch_plans(rv1, rv6, rv14, pt)
    n = len(li["l_shipdate"])
    tol = max(1e-12, n * 2.0 ** -53)
    st_f, st_g = check_plan(q1, [rv1], configs=(FUSED, GENERIC), rel_tol=tol, oracle_batch_rows=100_000)
    assert stat(st_f, "b200.fusedBatches") == 1 and stat(st_g, "b200.fusedBatches") == 0
    st_f, _ = check_plan(q6, [rv6], configs=(FUSED, GENERIC), rel_tol=tol, oracle_batch_rows=100_000)
    assert stat(st_f, "b200.fusedBatches") == 1
    st_f, _ = check_plan(q14, [rv14, pt], configs=(FUSED, GENERIC), rel_tol=tol, oracle_batch_rows=100_000)
    assert stat(st_f, "b200.fusedBatches") == 1
    check_plan(q1, [rv1], configs=(FUSED, GENERIC), batch_rows=7_000, rel_tol=tol, oracle_batch_rows=100_000)
