"""Function registries of the boundary: exec::registerVectorFunction with a
user-supplied device function, exec::registerAggregateFunction with a composed aggregate, and the
node-at-a-time VectorFunction::apply contract (velox/expression/VectorFunction.h:81-86,241;
velox/exec/Aggregate.h:525-575). The reference tests its registries with functions defined in the
test (velox/expression/tests/VectorFunctionTest / FunctionRegistryTest); so do these. The oracle has
no user functions, so expectations come from numpy — element-wise IEEE doubles, bit exact."""
import ctypes as C

import numpy as np
import pytest

from velox_b200._lib import lib
from velox_b200.plan import PlanBuilder, register_aggregate_function, register_scalar_function
from velox_b200.vector import BIGINT, BOOLEAN, DOUBLE, INTEGER, flat_vector, row_vector

HYPOT2 = "__device__ double hypot2(double a, double b) { return sqrt(a * a + b * b); }"
SQUARE = "__device__ double vb_square(double a) { return a * a; }"
ROOT = "__device__ double vb_root(double a) { return sqrt(a); }"
BUCKET = "__device__ long long bucket3(long long k, int m) { return (k % m + m) % m; }"
IS_BIG = "__device__ bool is_big(double a) { return a > 100.0; }"


def _register():
    register_scalar_function("hypot2", DOUBLE, [DOUBLE, DOUBLE], HYPOT2)
    register_scalar_function("square", DOUBLE, [DOUBLE], SQUARE, entry="vb_square")
    register_scalar_function("root", DOUBLE, [DOUBLE], ROOT, entry="vb_root")
    register_scalar_function("bucket3", BIGINT, [BIGINT, INTEGER], BUCKET)
    register_scalar_function("is_big", BOOLEAN, [DOUBLE], IS_BIG)
    register_aggregate_function("sum_sq", "sum", input_function="square")
    register_aggregate_function("rms", "avg", input_function="square", final_function="root")
    register_aggregate_function("total", "sum")  # a plain alias


def test_registered_functions_compile_into_expression_kernels():
    """No GPU: a plan calling registered functions goes through the expression compiler and every
    kernel of it is generated and NVRTC-compiled for sm_100a with the user's source spliced in."""
    _register()
    plan = (PlanBuilder().values(["a", "b", "k", "m"], [DOUBLE, DOUBLE, BIGINT, INTEGER])
            .filter("is_big(hypot2(a, b)) and bucket3(k, m) = 1")
            .project(["hypot2(a, b) * 2.0 as h", "bucket3(k, 3) + 1 as g", "square(a) as s"]).planNode())
    L = lib()
    progs, jit, total = C.c_int32(0), C.c_int32(0), C.c_int32(0)
    err = C.create_string_buffer(2048)
    rc = L.vb2_plan_jit_report(plan.sexpr.encode(), C.byref(progs), C.byref(jit), C.byref(total), err, 2048)
    assert rc == 0, err.value.decode()
    assert progs.value == 2 and jit.value == total.value == 2, (progs.value, jit.value, total.value)


def test_unknown_and_malformed_registrations_are_errors():
    _register()
    with pytest.raises(Exception):
        register_aggregate_function("bad1", "median")                       # not a device accumulator family
    with pytest.raises(Exception):
        register_aggregate_function("bad2", "sum", input_function="nope")  # transform must be registered
    with pytest.raises(ValueError):
        PlanBuilder().values(["a"], [DOUBLE]).project(["nope(a)"])
    with pytest.raises(ValueError):
        PlanBuilder().values(["a"], [DOUBLE]).singleAggregation([], ["nope(a)"])


def _table(n=20000, seed=5, nulls=True):
    rng = np.random.default_rng(seed)
    a = np.round(rng.normal(0, 80, n), 3)
    b = np.round(rng.normal(0, 80, n), 3)
    k = rng.integers(-1000, 1000, n)
    m = rng.integers(2, 9, n).astype(np.int32)
    an = rng.random(n) < 0.1 if nulls else np.zeros(n, bool)
    rv = row_vector(["a", "b", "k", "m"], [flat_vector(DOUBLE, [None if an[i] else float(a[i]) for i in range(n)]), flat_vector(DOUBLE, b),
                                          flat_vector(BIGINT, k), flat_vector(INTEGER, m)])
    return rv, a, b, k, m, an


@pytest.mark.gpu
def test_registered_scalar_functions_run_on_the_device():
    from velox_b200.task import run_plan
    _register()
    rv, a, b, k, m, an = _table()
    plan = (PlanBuilder().values(rv.names, rv.types)
            .project(["hypot2(a, b) as h", "bucket3(k, m) as g", "is_big(hypot2(a, b)) as big", "square(a) + 1.0 as s", "k"]).planNode())
    got, _ = run_plan(plan, [rv], batch_rows=4096)
    rows = got.rows()
    assert len(rows) == len(a)
    h = np.sqrt(a * a + b * b)
    g = np.mod(k, m)
    for i, r in enumerate(rows):
        assert r[4] == k[i]
        if an[i]:
            assert r[0] is None and r[2] is None and r[3] is None  # default NULL behaviour: NULL in -> NULL out
        else:
            assert r[0] == h[i] and r[2] == bool(h[i] > 100.0) and r[3] == a[i] * a[i] + 1.0, (i, r, h[i])
        assert r[1] == g[i]
    # as a filter, fused with built-in predicates
    f = PlanBuilder().values(rv.names, rv.types).filter("is_big(hypot2(a, b)) and bucket3(k, m) = 1").project(["k"]).planNode()
    got, _ = run_plan(f, [rv])
    keep = (~an) & (h > 100.0) & (g == 1)
    assert sorted(r[0] for r in got.rows()) == sorted(k[keep].tolist())


@pytest.mark.gpu
def test_registered_function_needs_the_jit():
    """The interpreter cannot run source text: with the JIT off the plan is an error, not a fallback."""
    from velox_b200._lib import VeloxRuntimeError
    from velox_b200.task import run_plan
    _register()
    rv, *_ = _table(1000)
    plan = PlanBuilder().values(rv.names, rv.types).project(["hypot2(a, b) as h"]).planNode()
    lib().vb2k_set_expression_jit(0)
    try:
        with pytest.raises(VeloxRuntimeError):
            run_plan(plan, [rv])
    finally:
        lib().vb2k_set_expression_jit(1)


@pytest.mark.gpu
@pytest.mark.parametrize("stages", ["single", "partial_final"])
def test_registered_aggregates(stages):
    from velox_b200.task import run_plan
    _register()
    rv, a, b, k, m, an = _table()
    aggs = ["sum_sq(a) as ss", "rms(b) as r", "total(k) as t", "count(0) as c", "rms(a) as ra"]
    pb = PlanBuilder().values(rv.names, rv.types).project(["m", "a", "b", "k"])
    plan = (pb.singleAggregation(["m"], aggs) if stages == "single" else pb.partialAggregation(["m"], aggs).localPartition([]).finalAggregation()).planNode()
    got, _ = run_plan(plan, [rv], batch_rows=3000)
    rows = {r[0]: r[1:] for r in got.rows()}
    assert set(rows) == set(np.unique(m).tolist())
    for g, (ss, r, t, c, ra) in rows.items():
        sel = m == g
        va = a[sel & ~an]
        assert c == int(sel.sum()) and t == int(k[sel].sum())
        assert abs(ss - float((va * va).sum())) <= 1e-12 * abs(ss)
        assert abs(r - float(np.sqrt((b[sel] ** 2).mean()))) <= 1e-12 * abs(r)
        assert abs(ra - float(np.sqrt((va ** 2).mean()))) <= 1e-12 * abs(ra)  # NULL inputs are skipped, as avg does


@pytest.mark.gpu
def test_vector_function_apply_contract():
    """VectorFunction::apply node-at-a-time: only the selected rows are computed and written; the rest
    of a pre-allocated result is preserved (VectorFunction.h:44-80). Built-in and user functions."""
    from velox_b200.vector import CColumn
    _register()
    L = lib()
    n = 1000
    rng = np.random.default_rng(11)
    a, b = rng.normal(0, 10, n), rng.normal(0, 10, n)
    sel = rng.random(n) < 0.5
    bits = np.zeros((n + 63) // 64, dtype=np.uint64)
    for i in np.nonzero(sel)[0]:
        bits[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    for name, want in (("hypot2", np.sqrt(a * a + b * b)), ("multiply", a * b), ("plus", a + b)):
        ca, cb = flat_vector(DOUBLE, a), flat_vector(DOUBLE, b)
        arr = (CColumn * 2)(ca.to_c(), cb.to_c())
        out = np.full(n, -7.0)
        nulls = np.zeros(n, dtype=np.uint8)
        err = C.create_string_buffer(1024)
        rc = L.vb2_scalar_function_apply(name.encode(), arr, 2, C.c_int64(n), bits.ctypes.data_as(C.c_void_p), DOUBLE, out.ctypes.data_as(C.c_void_p),
                                         nulls.ctypes.data_as(C.c_void_p), err, 1024)
        assert rc == 0, err.value.decode()
        assert np.array_equal(out[sel], want[sel]), name
        assert np.all(out[~sel] == -7.0), name  # untouched
        assert not nulls.any()
