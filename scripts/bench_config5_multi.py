"""BASELINE.json configs[4] across GPUs (run under torchrun) through the C++
operators: every rank holds 1/G of the (key, value) rows and runs the two-fragment plan

    values -> B200PartitionedOutput(hash(k) % G) -> B200Exchange -> B200HashAggregation(k: sum(v), count)

so rows travel once (16 B/row, peer-memory stores over NVLink or NCCL) and every rank aggregates its
own key slice; the result stays sharded by key, as the reference leaves it after a partitioned final
aggregation (velox/exec/tests/MultiFragmentTest.cpp aggregationMultiKey). Checks sum(count) == rows
and sum(sum(v)) against the closed form; --check compares every group with the CPU oracle's
single-process result over the union of the shards. Prints rows/s (max over ranks)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from bench_config5 import splitmix_keys
from velox_b200.comm import Comm
from velox_b200.kernels import flat_device
from velox_b200.plan import PlanBuilder
from velox_b200.task import Task
from velox_b200.vector import BIGINT, flat_vector, row_vector


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=float, default=1e9)
    ap.add_argument("--keys", type=float, default=1e8)
    ap.add_argument("--batch", type=float, default=2e9, help="rows per input batch (one batch per rank takes the exchange operator's fused single-key partition pass)")
    ap.add_argument("--iters", type=int, default=3)
    ap.add_argument("--check", action="store_true", help="compare every group with the CPU oracle (small sizes)")
    a = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm()
    rows, nkeys, batch = int(a.rows), int(a.keys), int(a.batch)
    r0, r1 = rows * rank // world, rows * (rank + 1) // world
    keys = splitmix_keys(r0, r1 - r0, nkeys)
    vals = torch.arange(r0, r1, device="cuda", dtype=torch.int64) % 1000
    plan = (PlanBuilder().values(["k", "v"], [BIGINT, BIGINT]).partitionedOutput(["k"])
            .singleAggregation(["k"], ["sum(v)", "count(0)"]).planNode())
    cfg = {"b200.result_on_device": "true"}
    # every rank feeds the same number of batches (an exchange round is collective)
    nb = max(1, -(-(rows // world + 1) // batch))
    times = []
    res = None
    for it in range(a.iters + 1):
        t = Task(plan, cfg)
        t.set_comm(comm)
        n = r1 - r0
        for b in range(nb):
            b0, b1 = n * b // nb, n * (b + 1) // nb
            t.add_input(0, [flat_device(BIGINT, keys[b0:b1]), flat_device(BIGINT, vals[b0:b1])])
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t0 = time.perf_counter()
        t._run_only()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device="cuda", dtype=torch.float64)
        dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        st = t.stats()
        (res,) = t.device_result()
        res = [c.clone() for c in res]
        t.close()
        if it or a.iters == 0:
            times.append(dt.item())
    tot = torch.tensor([int(res[2].sum().item()), int(res[1].sum().item()), res[0].numel()], device="cuda", dtype=torch.int64)
    dist.all_reduce(tot)
    want_sum = (rows // 1000) * (999 * 1000 // 2) + sum(range(rows % 1000))
    ok = tot[0].item() == rows and tot[1].item() == want_sum
    info = {}
    if a.check:
        mine = np.stack([c.cpu().numpy() for c in res], axis=1)  # [groups, 3] = key, sum, count
        parts = [None] * world if rank == 0 else None
        dist.gather_object(mine, parts, dst=0)
        if rank == 0:
            from oracle import pyoracle
            got = np.concatenate(parts)
            got = got[np.argsort(got[:, 0], kind="stable")]
            hk = splitmix_keys(0, rows, nkeys).cpu().numpy()
            hv = (np.arange(rows, dtype=np.int64) % 1000)
            rv = row_vector(["k", "v"], [flat_vector(BIGINT, hk), flat_vector(BIGINT, hv)])
            oplan = PlanBuilder().values(["k", "v"], [BIGINT, BIGINT]).singleAggregation(["k"], ["sum(v)", "count(0)"]).planNode()
            w = pyoracle.run_plan(oplan, [rv], threads=8)
            want = np.stack([np.asarray(c.values, dtype=np.int64) for c in w.columns], axis=1)
            want = want[np.argsort(want[:, 0], kind="stable")]
            same = got.shape == want.shape and bool((got == want).all())
            # every key on exactly one rank (the partition function routes by key)
            disjoint = len(np.unique(got[:, 0])) == got.shape[0]
            ok = ok and same and disjoint
            info = {"oracle_groups": int(want.shape[0]), "groups_equal_oracle": same, "keys_on_one_rank": disjoint}
    times.sort()
    sec = times[len(times) // 2]
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    if rank == 0:
        print(json.dumps({"world": world, "rows": rows, "distinct": int(tot[2].item()), "ok": bool(ok), "seconds": sec, "rows_per_s": rows / sec,
                          "batches_per_rank": nb, "peer_memory": comm.peer_memory, "exchanges": comm.exchanges(), **info,
                          "path": "values -> B200PartitionedOutput -> B200Exchange -> B200HashAggregation (C++ operators)",
                          "agg_mode_rank0": [v for k, v in st.items() if k.endswith("b200.aggMode")],
                          "slice_agg_rows_rank0": sum(v for k, v in st.items() if k.endswith("b200.sliceAggRows")),
                          "wall_ms_rank0": {k: round(v / 1e6, 2) for k, v in st.items() if k.endswith("WallNanos") and v > 1e5}}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if flag.item() else 1)


if __name__ == "__main__":
    main()
