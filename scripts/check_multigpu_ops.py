"""Multi-GPU parity of the OPERATOR path (run under torchrun): the multi-fragment Q1 / Q14 plans
(B200PartitionedOutput -> B200Exchange between the fragments, one process per GPU) over per-rank row
shards must equal the CPU oracle's single-process result over the union of the shards — counts and
keys exactly, floating-point columns within 1e-11 relative."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

import bench
from oracle import pyoracle
from velox_b200 import tpch
from velox_b200.comm import Comm


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm()
    rows, nparts = 300_007, 20_000
    li = tpch.gen_lineitem(rows, nparts, seed=42 + rank, device="cuda")
    part_all = tpch.gen_part(nparts, seed=43, device="cuda")
    p0, p1 = nparts * rank // world, nparts * (rank + 1) // world
    part = {k: v[p0:p1].contiguous() for k, v in part_all.items()}
    small = {k: v[:100].cpu() for k, v in li.items()}
    rv1s, rv14s, pts = bench.host_tables(small, {k: v[:100].cpu() for k, v in part_all.items()}, 100)
    d1, d14 = bench.plans_distributed(rv1s, rv14s, pts)
    c1, c14, cp = bench.device_inputs(li, part)
    outs = []
    for _ in range(2):  # twice: nothing may leak from one task into the next
        o1, s1 = bench.run_task(d1, [(0, c1)], comm=comm)
        o14, s14 = bench.run_task(d14, [(0, c14), (1, cp)], comm=comm)
        outs.append((o1, o14))
    sent = sum(v for k, v in s14.items() if k.endswith("b200.exchangeRowsSent"))
    recv = sum(v for k, v in s14.items() if k.endswith("b200.exchangeRowsReceived"))
    tot = torch.tensor([sent, recv], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
    ok = True
    info = {}
    if rank == 0:
        shards = [tpch.gen_lineitem(rows, nparts, seed=42 + r, device="cuda") for r in range(world)]
        full = {k: torch.cat([s[k] for s in shards]).cpu() for k in shards[0]}
        rv1, rv14, pt = bench.host_tables(full, {k: v.cpu() for k, v in part_all.items()}, rows * world)
        p1_, p14_ = bench.plans(rv1, rv14, pt)
        w1 = pyoracle.run_plan(p1_, [rv1], threads=8)
        w14 = pyoracle.run_plan(p14_, [rv14, pt], threads=8)
        for o1, o14 in outs:
            par = bench.parity(o1, o14, w1, w14)
            ok = ok and par["exact_columns_ok"] and par["fp_max_rel_err"] <= 1e-11
            info = par
        # conservation: rows sent == rows received over all ranks and exchanges
        ok = ok and int(tot[0]) == int(tot[1])
        fused = sum(v for k, v in s14.items() if k.endswith("b200.fusedBatches"))
        print(json.dumps({"ok": bool(ok), "world": world, "parity": info, "exchange_rows_sent": int(tot[0]), "exchange_rows_received": int(tot[1]),
                          "q14_fused_batches_rank0": fused, "q1_groups": o1.size,
                          "peer_memory": comm.peer_memory, "exchanges": comm.exchanges()}))
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
