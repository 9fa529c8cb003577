"""Arrow C data interface import in the C++ layer (velox/vector/arrow/Bridge.h:153-173,
tests modelled on velox/vector/arrow/tests/ArrowBridgeArrayTest.cpp: flat / NULLs / sliced arrays with an
offset / dictionary / strings / booleans). pyarrow exports the structs; vb2_task_add_arrow imports them."""
import ctypes as C
import gc

import numpy as np
import pyarrow as pa
import pytest

from velox_b200._lib import VeloxRuntimeError
from velox_b200.arrow import row_vector_from_arrow
from velox_b200.plan import PlanBuilder
from velox_b200.task import Task


def _table(n=5000, seed=1):
    rng = np.random.default_rng(seed)
    return pa.table({
        "k": pa.array(rng.integers(0, 50, n), type=pa.int64(), mask=rng.random(n) < 0.1),
        "d": pa.array(rng.integers(8000, 11000, n).astype(np.int32), type=pa.date32()),
        "x": pa.array(np.round(rng.normal(0, 10, n), 2), type=pa.float64(), mask=rng.random(n) < 0.05),
        "b": pa.array(rng.random(n) < 0.5, type=pa.bool_(), mask=rng.random(n) < 0.1),
        "s": pa.array(rng.choice(["x", "yy", "", "promo long string beyond twelve bytes"], n).tolist(), type=pa.string()),
        "c": pa.DictionaryArray.from_arrays(pa.array(rng.integers(0, 3, n).astype(np.int32), mask=rng.random(n) < 0.1), pa.array(["A", "N", "R"])),
    })


def test_import_consumes_and_releases_the_arrow_structs():
    """No GPU: the import succeeds, ownership moves to the task (pyarrow's export is released when the
    task is freed), unsupported formats are errors."""
    t = _table(1000)
    rv = row_vector_from_arrow(t)
    plan = PlanBuilder().values(rv.names, rv.types).planNode()
    before = pa.total_allocated_bytes()
    task = Task(plan)
    batch = t.combine_chunks().to_batches()[0]
    task.add_arrow(0, batch)
    task.add_arrow(0, batch.slice(100, 300))  # a sliced batch: children carry an offset
    del batch
    task.close()
    del t
    gc.collect()
    assert pa.total_allocated_bytes() < before  # the task released its references to the exported buffers
    bad = pa.record_batch([pa.array([1.5, 2.5], type=pa.float32())], names=["f"])
    task = Task(PlanBuilder().values(["f"], [6]).planNode())
    with pytest.raises(VeloxRuntimeError):
        task.add_arrow(0, bad)
    task.close()


@pytest.mark.gpu
def test_arrow_batches_feed_the_operators():
    """The same plan over batches imported through the Arrow C data interface and over the same data
    handed in as host columns: identical results (row for row after ORDER BY)."""
    from velox_b200.task import run_plan
    t = _table(20000)
    rv = row_vector_from_arrow(t)
    plan = (PlanBuilder().values(rv.names, rv.types).filter("x > -5.0 and (b or k < 25) and s <> 'x'")
            .project(["k", "c", "x * 2.0 as y", "d"]).singleAggregation(["c", "k"], ["sum(y)", "count(0)", "max(d)"]).orderBy(["c", "k"]).planNode())
    want, _ = run_plan(plan, [rv])
    task = Task(plan)
    try:
        whole = t.combine_chunks().to_batches()[0]
        for off in range(0, whole.num_rows, 6000):  # slices: every child array has offset != 0 after the first
            task.add_arrow(0, whole.slice(off, 6000))
        got = task.run()
    finally:
        task.close()
    # same keys in the same (sorted) order; FP sums within the usual tolerance (the batch boundaries differ)
    assert [r[:2] + r[3:] for r in got.rows()] == [r[:2] + r[3:] for r in want.rows()] and len(got.rows()) > 100
    assert all(abs(a[2] - b[2]) <= 1e-12 * max(1.0, abs(b[2])) for a, b in zip(got.rows(), want.rows()))
