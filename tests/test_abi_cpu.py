"""CPU-side checks of the drop-in boundary (no GPU, no compute calls):
  * libvelox_b200.so loads and exports every symbol include/*.h declares;
  * the plan front-end produces the plan text both libraries read; the product library's own
    reader accepts it (vb2_task_create) and rejects malformed text with VeloxRuntimeError;
  * without a GPU the product path fails loudly — there is no CPU fallback."""
import torch  # noqa: F401  (first: our library must bind to the libnccl torch bundles)
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "velox_b200", "lib", "libvelox_b200.so")


def declared_symbols():
    names = set()
    for h in ("velox_b200.h", "velox_b200_kernels.h"):
        text = open(os.path.join(ROOT, "include", h)).read()
        text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
        names |= set(re.findall(r"\b(vb2k?_[a-z0-9_]+)\s*\(", text))
    return names


def test_library_exports_every_declared_symbol():
    assert os.path.exists(LIB), "run `make` / __graft_entry__.build() first"
    lib = C.CDLL(LIB)
    missing = [s for s in sorted(declared_symbols()) if not hasattr(lib, s)]
    assert not missing, f"declared in include/*.h but not exported: {missing}"
    assert len(declared_symbols()) > 40


def test_fused_registry_and_signatures():
    lib = C.CDLL(LIB)
    lib.vb2k_fused_signature.restype = C.c_char_p
    sigs = [lib.vb2k_fused_signature(i).decode() for i in range(lib.vb2k_fused_count())]
    from velox_b200 import tpch
    from velox_b200.queries import Q14_PROBE_SIG, Q14_SCAN_SIG
    for s in (tpch.Q1_SIG, tpch.Q6_SIG, tpch.Q14_SIG, Q14_SCAN_SIG, Q14_PROBE_SIG):
        assert s in sigs and lib.vb2k_fused_find(s.encode()) == sigs.index(s)
    assert lib.vb2k_fused_find(b"F:true;P:nope") == -1


def test_plan_text_roundtrip_and_errors():
    from velox_b200.plan import PlanBuilder
    from velox_b200.vector import BIGINT, DOUBLE, INTEGER, VARCHAR
    lib = C.CDLL(LIB)
    lib.vb2_task_create.restype = C.c_void_p
    lib.vb2_task_create.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int32]
    lib.vb2_task_free.argtypes = [C.c_void_p]
    names, types = ["k", "v", "d", "s"], [BIGINT, DOUBLE, INTEGER, VARCHAR]
    build = PlanBuilder().values(["bk", "bs"], [BIGINT, VARCHAR], source=1)
    plan = (PlanBuilder().values(names, types).filter("d between '1995-01-01'::DATE and '1995-12-31'::DATE and v * 2.0 > 1.5 or s like 'A%'")
            .project(["k", "case when v > 0.0 then v else -v end as a", "cast(d as bigint) + k as e", "s"])
            .hashJoin(["k"], ["bk"], build, "a > 1.0", ["k", "a", "bs"])
            .partialAggregation(["bs"], ["sum(a)", "avg(a)", "count(0)", "min(k)"]).finalAggregation().planNode())
    assert plan.sexpr.startswith("(aggregation final")
    err = C.create_string_buffer(1024)
    h = lib.vb2_task_create(plan.sexpr.encode(), b"b200.fused_pipelines=false", err, 1024)
    assert h, err.value
    lib.vb2_task_free(h)
    # DISTINCT aggregates travel as (fn col (distinct)); the front end only takes them over a column of a single aggregation
    dplan = PlanBuilder().values(names, types).singleAggregation(["k"], ["count(distinct d)", "sum(DISTINCT d) as s"]).planNode()
    assert "(count 2 (distinct))" in dplan.sexpr and "(sum 2 (distinct))" in dplan.sexpr
    h = lib.vb2_task_create(dplan.sexpr.encode(), b"", err, 1024)
    assert h, err.value
    lib.vb2_task_free(h)
    for bad_agg in ("count(distinct 0)", "count(distinct *)"):
        with pytest.raises(ValueError):
            PlanBuilder().values(names, types).singleAggregation(["k"], [bad_agg])
    with pytest.raises(ValueError):
        PlanBuilder().values(names, types).partialAggregation(["k"], ["count(distinct d)"])
    for bad in (b"(filter (lt (field 9) (i64 1)) (values 0 (BIGINT)))", b"(values 0 (BIGINT)", b"(frobnicate)", b"(project ((nosuchfn (field 0))) (values 0 (BIGINT)))",
                b"(aggregation single (keys 0) (aggs (count 0 (unique))) (values 0 (BIGINT)))"):
        assert not lib.vb2_task_create(bad, b"", err, 1024)
        assert err.value.startswith(b"VeloxRuntimeError")


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from velox_b200._lib import VeloxRuntimeError
    from velox_b200.plan import PlanBuilder
    from velox_b200.task import run_plan
    from velox_b200.vector import BIGINT, flat_vector, row_vector
    rv = row_vector(["a"], [flat_vector(BIGINT, [1, 2, 3])])
    with pytest.raises(VeloxRuntimeError, match="CUDA"):
        run_plan(PlanBuilder().values(rv.names, rv.types).filter("a > 1").planNode(), [rv])


def test_sql_frontend_types_and_literals():
    from velox_b200.plan import parse_expr
    from velox_b200.vector import BIGINT, BOOLEAN, DOUBLE, INTEGER
    names, types = ["q", "d", "k", "i"], [DOUBLE, INTEGER, BIGINT, INTEGER]
    assert parse_expr("q < 24", names, types)[0].sexpr == "(lt (field 0) (f64 24.0))"       # literal adopts the column's type
    assert parse_expr("d <= '1998-09-02'::DATE", names, types)[0].sexpr == "(lte (field 1) (i32 10471))"
    assert parse_expr("k + i", names, types)[0].sexpr == "(plus (field 2) (cast BIGINT (field 3)))"
    assert parse_expr("q * (1.0 - q) as x", names, types) [1] == "x"
    assert parse_expr("not (k = 1 or q is null)", names, types)[0].type == BOOLEAN
