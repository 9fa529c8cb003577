"""Phase timing of the hash-partitioned Q14 (run under torchrun): each phase bracketed by a device
synchronisation, so the numbers are serial costs (the product path overlaps scan and build side)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from velox_b200 import tpch
from velox_b200.comm import Comm
from velox_b200.kernels import (FusedScanAgg, FusedScanCompact, flat_device, gather, hash_columns, partition_ids,
                                partition_scatter_order)
from velox_b200.queries import Q14, Q14_SCAN_SIG
from velox_b200.vector import BIGINT


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    comm = Comm()
    sf = float(os.environ.get("SF", "100"))
    rows_total, nparts = int(6_000_379.02 * sf), int(200_000 * sf)
    r0, r1 = rows_total * rank // world, rows_total * (rank + 1) // world
    rows = r1 - r0
    li = tpch.gen_lineitem(rows, nparts, seed=42 + rank, device="cuda")
    part_all = tpch.gen_part(nparts, seed=43, device="cuda")
    p0, p1 = nparts * rank // world, nparts * (rank + 1) // world
    part = {k: v[p0:p1].contiguous() for k, v in part_all.items()}
    del part_all
    q = Q14(comm)
    for _ in range(3):
        q.launch(li, part, rows); q.merge()
    torch.cuda.synchronize()

    T = {}

    def phase(name, fn, iters=10):
        torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(iters):
            out = fn()
        torch.cuda.synchronize()
        T[name] = round((time.perf_counter() - t) / iters * 1e3, 4)
        return out

    scan = q.scan
    phase("scan_compact", lambda: scan.run([li["l_shipdate"], li["l_partkey"], li["l_extendedprice"], li["l_discount"]], rows,
                                           pf=[1.0], pi=[tpch.Q14_SHIP_LO, tpch.Q14_SHIP_HI]))
    n, (lk, rev) = phase("scan_result_sync", lambda: scan.result([torch.int64, torch.float64]))
    T["compact_rows"] = int(n)

    def part_prep(key, payload):
        h = hash_columns([flat_device(BIGINT, key)])
        ids = partition_ids(h, world)
        counts, order = partition_scatter_order(ids, world)
        return counts, gather(key, order), gather(payload, order)

    counts, sk, sp = phase("li_hash_partition_gather", lambda: part_prep(lk, rev))
    sc, rc = phase("li_counts_exchange", lambda: comm.exchange_counts_dev(counts))
    rk, rrev = phase("li_all_to_all", lambda: comm.all_to_all_columns([sk, sp], sc, rc))
    pcounts, pk_s, pt_s = phase("part_hash_partition_gather", lambda: part_prep(part["p_partkey"], part["p_type"]))
    psc, prc = phase("part_counts_exchange", lambda: comm.exchange_counts_dev(pcounts))
    pk, pt = phase("part_all_to_all", lambda: comm.all_to_all_columns([pk_s, pt_s], psc, prc))
    slot_flags, join_min = phase("build", lambda: q._build(pk, pt))

    def probe():
        q.probe.reset()
        q.probe.add_batch([rk, rrev], rk.numel(), pf=[0.0], join={"slot_flags": slot_flags, "min": join_min})

    phase("probe", probe)
    phase("merge_allreduce", q.merge)
    phase("whole_launch_merge", lambda: (q.launch(li, part, rows), q.merge()))
    if rank == 0:
        print(json.dumps({"world": world, "rows_per_rank": rows, "ms": T}))
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
