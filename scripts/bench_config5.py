"""BASELINE.json configs[4]: high-cardinality HashAggregation — N BIGINT rows, D distinct keys —
through the operator-level C ABI with device-resident input (Task -> B200HashAggregation, hash mode).
Reports rows/s and the streaming-roofline fraction (algorithmic bytes = 16 B/row in + 16 B/group out)."""
import argparse
import json
import os
import sys
import time

os.environ.setdefault("VB2_SYNC_TIMING", "1")  # operator wall-time stats include their kernels (read at library load)

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from velox_b200.kernels import flat_device
from velox_b200.plan import PlanBuilder
from velox_b200.task import Task
from velox_b200.vector import BIGINT


def _lsr(x, k):
    return (x >> k) & ((1 << (64 - k)) - 1)


def splitmix_keys(start: int, count: int, nkeys: int, chunk: int = 1 << 27) -> torch.Tensor:
    """keys = splitmix64(i) % nkeys (as unsigned), i in [start, start + count)."""
    out = torch.empty(count, dtype=torch.int64, device="cuda")
    def c(v):  # two's-complement constant
        return v - (1 << 64) if v >= (1 << 63) else v
    for c0 in range(0, count, chunk):
        n = min(chunk, count - c0)
        z = (torch.arange(start + c0, start + c0 + n, device="cuda", dtype=torch.int64) + 1) * c(0x9E3779B97F4A7C15)
        z = (z ^ _lsr(z, 30)) * c(0xBF58476D1CE4E5B9)
        z = (z ^ _lsr(z, 27)) * c(0x94D049BB133111EB)
        z = z ^ _lsr(z, 31)
        r = torch.remainder(z, nkeys)
        r = torch.where(z < 0, torch.remainder(r + ((1 << 64) % nkeys), nkeys), r)
        out[c0:c0 + n] = r
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=float, default=1e9)
    ap.add_argument("--keys", type=float, default=1e8)
    ap.add_argument("--batch", type=float, default=2.5e8)
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--radix", action="store_true", help="radix-partition the batches first (b200.agg_radix_partition)")
    ap.add_argument("--table", action="store_true", help="global-table path only (b200.agg_slice_aggregation=false)")
    a = ap.parse_args()
    rows, nkeys, batch = int(a.rows), int(a.keys), int(a.batch)
    keys = splitmix_keys(0, rows, nkeys)
    vals = torch.arange(rows, device="cuda", dtype=torch.int64) % 1000
    torch.cuda.synchronize()
    plan = PlanBuilder().values(["k", "v"], [BIGINT, BIGINT]).singleAggregation(["k"], ["sum(v)", "count(0)"]).planNode()
    times = []
    cfg = {"b200.result_on_device": "true"}  # 100 M result groups = 2.4 GB: the consumer of such a result sits on the device
    if a.radix:
        cfg["b200.agg_radix_partition"] = "true"
    if a.table:
        cfg["b200.agg_slice_aggregation"] = "false"
    for it in range(a.iters + 1):
        t = Task(plan, cfg)
        for r0 in range(0, rows, batch):
            r1 = min(rows, r0 + batch)
            t.add_input(0, [flat_device(BIGINT, keys[r0:r1]), flat_device(BIGINT, vals[r0:r1])])
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        t._run_only()
        dt = time.perf_counter() - t0
        st = t.stats()
        (res,) = t.device_result()
        groups = res[0].numel()
        total = int(res[2].sum().item())
        ssum = int(res[1].sum().item())
        t.close()
        if it or a.iters == 0:
            times.append(dt)
    assert total == rows and ssum == int(vals.sum().item()), (total, ssum)
    sec = sorted(times)[len(times) // 2]
    bytes_alg = rows * 16 + groups * 16
    # device pipeline = the aggregation operator (addInput of every batch + getOutput), kernels included
    dev = sum(v for k, v in st.items() if "B200HashAggregation" in k and k.endswith("WallNanos")) / 1e9
    sector_bytes = rows * (16 + 64) + groups * 16  # + one 32-B group-row sector read and written per input row
    print(json.dumps({"rows": rows, "distinct": groups, "device_pipeline_seconds": dev, "rows_per_s": rows / dev,
                      "algorithmic_GBps": bytes_alg / dev / 1e9, "frac_of_measured_hbm": bytes_alg / dev / 1e9 / 6570.9,
                      "with_sector_rmw_GBps": sector_bytes / dev / 1e9, "with_sector_rmw_frac": sector_bytes / dev / 1e9 / 6570.9,
                      "task_seconds_operator_api_result_on_device": sec, "api_over_device": sec / dev,
                      "agg_mode": [v for k, v in st.items() if k.endswith("b200.aggMode")],
                      "slice_agg_rows": sum(v for k, v in st.items() if k.endswith("b200.sliceAggRows")),
                      "wall_ms": {k: round(v / 1e6, 2) for k, v in st.items() if k.endswith("WallNanos") and v > 1e5}}))


if __name__ == "__main__":
    main()
