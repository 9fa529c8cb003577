"""BASELINE.json configs[4]: high-cardinality HashAggregation — N BIGINT rows, D distinct keys —
through the operator-level C ABI with device-resident input (Task -> B200HashAggregation, hash mode).
Reports rows/s and the streaming-roofline fraction (algorithmic bytes = 16 B/row in + 16 B/group out)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from velox_b200.kernels import flat_device
from velox_b200.plan import PlanBuilder
from velox_b200.task import Task
from velox_b200.vector import BIGINT


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=float, default=1e9)
    ap.add_argument("--keys", type=float, default=1e8)
    ap.add_argument("--batch", type=float, default=2.5e8)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    rows, nkeys, batch = int(a.rows), int(a.keys), int(a.batch)
    g = torch.Generator(device="cuda")
    g.manual_seed(7)
    keys = torch.randint(0, nkeys, (rows,), generator=g, device="cuda", dtype=torch.int64)
    vals = torch.arange(rows, device="cuda", dtype=torch.int64) % 1000
    torch.cuda.synchronize()
    plan = PlanBuilder().values(["k", "v"], [BIGINT, BIGINT]).singleAggregation(["k"], ["sum(v)", "count(0)"]).planNode()
    times = []
    for it in range(a.iters + 1):
        t = Task(plan)
        for r0 in range(0, rows, batch):
            r1 = min(rows, r0 + batch)
            t.add_input(0, [flat_device(BIGINT, keys[r0:r1]), flat_device(BIGINT, vals[r0:r1])])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = t.run()
        dt = time.perf_counter() - t0
        st = t.stats()
        t.close()
        if it:
            times.append(dt)
    groups = out.size
    total = int(out.columns[2].values.sum())
    ssum = int(out.columns[1].values.sum())
    assert total == rows and ssum == int(vals.sum().item()), (total, ssum)
    sec = sorted(times)[len(times) // 2]
    bytes_alg = rows * 16 + groups * 16
    print(json.dumps({"rows": rows, "distinct": groups, "seconds": sec, "rows_per_s": rows / sec, "algorithmic_GBps": bytes_alg / sec / 1e9,
                      "frac_of_measured_hbm": bytes_alg / sec / 1e9 / 6570.9, "includes": "result device->host copy of all groups",
                      "agg_mode": [v for k, v in st.items() if k.endswith("b200.aggMode")]}))


if __name__ == "__main__":
    main()
