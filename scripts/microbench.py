"""Kernel-level timing of the fused pipelines on device-resident synthetic TPC-H columns.
Development aid (bench.py is the contract harness)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from velox_b200 import tpch
from velox_b200.kernels import FusedScanAgg


def time_it(fn, warmup=3, iters=10):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        e.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=10)
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    rows = int(tpch.LINEITEM_ROWS_PER_SF * a.sf)
    nparts = int(tpch.PART_ROWS_PER_SF * a.sf)
    t0 = time.time()
    li = tpch.gen_lineitem(rows, nparts, device="cuda")
    part = tpch.gen_part(nparts, device="cuda")
    torch.cuda.synchronize()
    print(f"generated {rows} rows in {time.time() - t0:.1f}s", flush=True)
    out = {"sf": a.sf, "rows": rows}

    q6 = FusedScanAgg(tpch.Q6_SIG)
    def run_q6():
        q6.reset()
        q6.add_batch([li["l_shipdate"], li["l_discount"], li["l_quantity"], li["l_extendedprice"]], rows,
                     pf=[0.05, 0.07, 24.0], pi=[tpch.Q6_SHIP_LO, tpch.Q6_SHIP_HI])
    med, best = time_it(run_q6, iters=a.iters) if a.only in ('', 'q6') else (1.0, 1.0)
    out["q6"] = {"ms": med, "best_ms": best, "GBps": rows * tpch.Q6_BYTES_PER_ROW / med / 1e6, "rows_per_s": rows / med * 1e3,
                 "sum": q6.sums.item(), "count": q6.counts.item()}

    q1 = FusedScanAgg(tpch.Q1_SIG, ngroups=6)
    def run_q1():
        q1.reset()
        q1.add_batch([li["l_shipdate"], li["l_quantity"], li["l_extendedprice"], li["l_discount"], li["l_tax"]], rows,
                     pf=[1.0, 1.0, 1.0], pi=[tpch.Q1_SHIPDATE_LT], keys=[li["l_returnflag"], li["l_linestatus"]],
                     key_min=[0, 0], key_mult=[2, 1])
    med, best = time_it(run_q1, iters=a.iters) if a.only in ('', 'q1') else (1.0, 1.0)
    out["q1"] = {"ms": med, "best_ms": best, "GBps": rows * tpch.Q1_BYTES_PER_ROW / med / 1e6, "rows_per_s": rows / med * 1e3,
                 "counts": q1.counts.tolist()}

    head = torch.zeros(nparts, dtype=torch.int32, device="cuda")
    head[part["p_partkey"] - 1] = torch.arange(1, nparts + 1, dtype=torch.int32, device="cuda")
    flag = torch.tensor([1 if s.startswith("PROMO") else 0 for s in tpch.PTYPE_DICT], dtype=torch.uint8, device="cuda")
    from velox_b200.kernels import join_slot_flags
    slot_flags = join_slot_flags(head, part["p_type"], flag)
    q14 = FusedScanAgg(tpch.Q14_SIG)
    def run_q14():
        q14.reset()
        q14.add_batch([li["l_shipdate"], li["l_partkey"], li["l_extendedprice"], li["l_discount"]], rows,
                      pf=[1.0, 1.0, 0.0], pi=[tpch.Q14_SHIP_LO, tpch.Q14_SHIP_HI],
                      join={"slot_flags": slot_flags, "min": 1})
    med, best = time_it(run_q14, iters=a.iters) if a.only in ('', 'q14') else (1.0, 1.0)
    s = q14.sums.tolist()
    out["q14_probe"] = {"ms": med, "best_ms": best, "GBps": rows * tpch.Q14_BYTES_PER_ROW / med / 1e6, "rows_per_s": rows / med * 1e3,
                        "promo_revenue": (100.0 * s[1] / s[0]) if s[0] else None, "count": q14.counts.item()}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
