"""BASELINE.json configs[4] across GPUs (run under torchrun): every rank holds
1/G of the (key, value) rows, partitions them by VectorHasher-hash(key) % G, ONE all-to-all of
16 B/row, then the local final aggregation (Task -> B200HashAggregation, hash mode) over its key
slice. Checks sum(count) == rows and sum(sum(v)) against the closed form; prints rows/s (max over ranks)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from bench_config5 import splitmix_keys
from velox_b200.comm import Comm
from velox_b200.kernels import flat_device, gather, hash_columns, partition_ids, partition_scatter_order
from velox_b200.plan import PlanBuilder
from velox_b200.task import Task
from velox_b200.vector import BIGINT


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=float, default=1e9)
    ap.add_argument("--keys", type=float, default=1e8)
    ap.add_argument("--iters", type=int, default=3)
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm()
    rows, nkeys = int(a.rows), int(a.keys)
    r0, r1 = rows * rank // world, rows * (rank + 1) // world
    keys = splitmix_keys(r0, r1 - r0, nkeys)
    vals = torch.arange(r0, r1, device="cuda", dtype=torch.int64) % 1000
    plan = PlanBuilder().values(["k", "v"], [BIGINT, BIGINT]).singleAggregation(["k"], ["sum(v)", "count(0)"]).planNode()
    times, phases = [], {}
    for it in range(a.iters + 1):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        h = hash_columns([flat_device(BIGINT, keys)])
        ids = partition_ids(h, world)
        counts, order = partition_scatter_order(ids, world)
        sk, sv = gather(keys, order), gather(vals, order)
        sc, rc = comm.exchange_counts_dev(counts)
        rk, rv = comm.all_to_all_columns([sk, sv], sc, rc)
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        t = Task(plan)
        t.add_input(0, [flat_device(BIGINT, rk), flat_device(BIGINT, rv)])
        out = t.run()
        t.close()
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        dt = torch.tensor([t2 - t0, t1 - t0, t2 - t1], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        if it:
            times.append(dt.tolist())
        del h, ids, order, sk, sv, rk, rv
    tot = torch.tensor([int(out.columns[2].values.sum()), int(out.columns[1].values.sum()), out.size], device="cuda", dtype=torch.int64)
    dist.all_reduce(tot)
    want_sum = (rows // 1000) * (999 * 1000 // 2) + sum(range(rows % 1000))
    ok = tot[0].item() == rows and tot[1].item() == want_sum
    times.sort()
    sec, ex, agg = times[len(times) // 2]
    if rank == 0:
        print(json.dumps({"world": world, "rows": rows, "distinct": int(tot[2].item()), "ok": bool(ok), "seconds": sec, "exchange_s": ex, "aggregate_s": agg,
                          "rows_per_s": rows / sec}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
