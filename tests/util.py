"""Shared helpers of the parity tests: run one plan through the product (C ABI -> B200 operators)
and through the CPU oracle, compare as multisets the way the reference's assertQuery does
(velox/exec/tests/utils/QueryAssertions.h:36,305-325): rows are matched on their non-floating
columns, floating columns within a tolerance; NaN equals NaN."""
import math

import numpy as np

from oracle import pyoracle
from velox_b200._lib import VeloxUserError
from velox_b200.task import run_plan


def _sort_key(row):
    out = []
    for v in row:
        if v is None:
            out.append((0, 0))
        elif isinstance(v, float):
            out.append((1, 0.0 if math.isnan(v) else float(f"{v:.9e}")) if not math.isnan(v) else (2, 0.0))
        elif isinstance(v, str):
            out.append((3, v))
        else:
            out.append((1, v))
    return tuple(out)


def assert_equal_results(got, want, rel_tol=1e-12, abs_tol=0.0):
    g, w = got.rows(), want.rows()
    assert len(g) == len(w), f"row counts differ: got {len(g)}, want {len(w)}\n got={g[:5]}\nwant={w[:5]}"
    g, w = sorted(g, key=_sort_key), sorted(w, key=_sort_key)
    for a, b in zip(g, w):
        assert len(a) == len(b)
        for x, y in zip(a, b):
            if x is None or y is None:
                assert x is None and y is None, (a, b)
            elif isinstance(y, float):
                if math.isnan(y):
                    assert math.isnan(x), (a, b)
                elif math.isinf(y):
                    assert x == y, (a, b)
                else:
                    assert abs(x - y) <= max(abs_tol, rel_tol * abs(y)), (a, b)
            else:
                assert x == y, (a, b)  # integers, booleans, strings, counts: bit exact


def check_plan(plan, sources, configs=({},), batch_rows=None, rel_tol=1e-12, oracle_batch_rows=10000):
    """Runs the plan on the oracle and, for every config, on the product; returns product stats. Integer / key / boolean /
    string columns must be equal; DOUBLE columns within rel_tol (the general aggregation paths add with atomics, so the
    order of a floating-point sum is not fixed from run to run — only the fused kernels are bit-reproducible)."""
    want = pyoracle.run_plan(plan, sources, threads=1, batch_rows=oracle_batch_rows)
    stats = []
    for cfg in configs:
        got, st = run_plan(plan, sources, config=cfg, batch_rows=batch_rows)
        assert_equal_results(got, want, rel_tol=rel_tol)
        stats.append(st)
    return stats


def check_user_error(plan, sources, configs=({},)):
    """Both sides must raise the user-error class (VeloxUserError)."""
    try:
        pyoracle.run_plan(plan, sources)
        raise AssertionError("oracle did not raise")
    except pyoracle.OracleUserError:
        pass
    for cfg in configs:
        try:
            run_plan(plan, sources, config=cfg)
            raise AssertionError("product did not raise")
        except VeloxUserError:
            pass


def stat(stats, suffix):
    return sum(v for k, v in stats.items() if k.endswith(suffix))


GENERIC = {"b200.fused_pipelines": "false"}
FUSED = {}
