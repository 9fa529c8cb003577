"""Plan shapes OUTSIDE the ahead-of-time fused registry at SF100 through the Task API (device-resident
input): the pipeline JIT (csrc/fused_jit.cu) instantiates the fused-scan templates for them at run time.
Reports the fused kernel's device time (operator stat b200.fusedScanNanos, CUDA events on the launch
stream), rows/s and the HBM roofline fraction on algorithmic bytes (each referenced column once)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from velox_b200 import tpch
from velox_b200.kernels import flat_device
from velox_b200.plan import PlanBuilder
from velox_b200.task import Task
from velox_b200.vector import DOUBLE, INTEGER


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=100)
    a = ap.parse_args()
    rows = int(tpch.LINEITEM_ROWS_PER_SF * a.sf)
    li = tpch.gen_lineitem(rows, int(tpch.PART_ROWS_PER_SF * a.sf), device="cuda")
    names = ["l_shipdate", "l_quantity", "l_extendedprice", "l_discount", "l_tax"]
    types = [INTEGER, DOUBLE, DOUBLE, DOUBLE, DOUBLE]
    cols = [flat_device(INTEGER, li["l_shipdate"])] + [flat_device(DOUBLE, li[c]) for c in names[1:]]
    peak, _ = bench.peaks()
    shapes = {
        # Q6 with a fourth predicate: 4 + 8 * 4 = 36 B/row
        "q6_plus_tax_predicate": (PlanBuilder().values(names, types)
                                  .filter("l_shipdate between '1994-01-01'::DATE and '1994-12-31'::DATE and l_discount between 0.05 and 0.07 "
                                          "and l_quantity < 24.0 and l_tax < 0.07")
                                  .project(["l_extendedprice * l_discount"]).singleAggregation([], ["sum(p0)"]).planNode(), 36),
        # a global Q1-like aggregation with different expressions: 4 + 8 * 4 = 36 B/row
        "global_q1_variant": (PlanBuilder().values(names, types).filter("l_shipdate < '1998-09-03'::DATE and l_quantity < 45.0")
                              .project(["l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS c", "l_extendedprice / (1.0 + l_discount) AS d", "l_quantity"])
                              .singleAggregation([], ["sum(c)", "sum(d)", "sum(l_quantity)", "count(0)"]).planNode(), 36),
        # very selective filter: the late-materialisation pair (filter bitmap + gather-aggregate); 4 B/row + survivors
        "selective_filter": (PlanBuilder().values(names, types).filter("l_shipdate between '1995-09-01'::DATE and '1995-09-03'::DATE")
                             .project(["l_extendedprice * (1.0 - l_discount) AS r"]).singleAggregation([], ["sum(r)", "count(0)"]).planNode(), 4),
    }
    out = {"sf": a.sf, "rows": rows, "peak_GBps": peak}
    for name, (plan, bpr) in shapes.items():
        ts, st = [], {}
        for it in range(5):
            t = Task(plan)
            t.add_input(0, cols)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = t.run()
            ts.append(time.perf_counter() - t0)
            st = t.stats()
            t.close()
        ms = sorted(ts[1:])[len(ts[1:]) // 2] * 1e3
        fused_ns = sum(v for k, v in st.items() if k.endswith("b200.fusedScanNanos"))
        kern_ms = fused_ns / 1e6 if fused_ns else None
        out[name] = {"task_ms": ms, "fused_kernel_ms": kern_ms, "algorithmic_bytes_per_row": bpr,
                     "algorithmic_GBps_task": rows * bpr / ms / 1e6, "frac_of_hbm_peak_task": rows * bpr / ms / 1e6 / peak,
                     "algorithmic_GBps_kernel": rows * bpr / kern_ms / 1e6 if kern_ms else None,
                     "frac_of_hbm_peak_kernel": rows * bpr / kern_ms / 1e6 / peak if kern_ms else None,
                     "fused_batches": sum(v for k, v in st.items() if k.endswith("b200.fusedBatches")),
                     "selective_batches": sum(v for k, v in st.items() if k.endswith("b200.selectiveBatches")),
                     "generic_batches": sum(v for k, v in st.items() if k.endswith("b200.genericBatches")),
                     "result": [x for x in r.rows()[0]]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
