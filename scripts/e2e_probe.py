"""Where does the end-to-end (host columns -> result) time go? (1) raw pinned host->device bandwidth
from the NUMA node the process happens to run on and from the GPU's own node, (2) one e2e step of
bench.py at a reduced scale factor with the operators' wall-time stats."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402


def gpu_numa():
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(0)
    bus = pynvml.nvmlDeviceGetPciInfo(h).busId
    bus = bus.decode() if isinstance(bus, bytes) else bus
    short = bus.lower()[-12:]  # 0000:xx:yy.z
    node, cpus = None, None
    p = f"/sys/bus/pci/devices/{short}"
    try:
        node = int(open(p + "/numa_node").read())
        cpus = open(p + "/local_cpulist").read().strip()
    except Exception as e:
        node = f"unreadable: {e}"
    return bus, node, cpus


def parse_cpulist(s):
    out = []
    for part in s.split(","):
        if "-" in part:
            a, b = part.split("-")
            out += list(range(int(a), int(b) + 1))
        elif part:
            out.append(int(part))
    return out


def h2d_gbs(nbytes=4 << 30, reps=3):
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes * reps / (time.perf_counter() - t0) / 1e9


def main():
    out = {}
    bus, node, cpus = gpu_numa()
    out["gpu"] = {"bus": bus, "numa_node": node, "local_cpulist": cpus, "host_cpus": os.cpu_count(), "affinity_before": len(os.sched_getaffinity(0))}
    torch.cuda.set_device(0)
    out["h2d_GBps_default_affinity"] = h2d_gbs()
    if cpus:
        os.sched_setaffinity(0, parse_cpulist(cpus))
        out["h2d_GBps_gpu_local_affinity"] = h2d_gbs()
    # one e2e step at SF30 with operator stats
    import bench
    from velox_b200 import tpch
    from velox_b200.task import Task, UploadCache, split_rowvector
    sf = float(os.environ.get("PROBE_SF", "30"))
    rows = int(tpch.LINEITEM_ROWS_PER_SF * sf)
    nparts = int(tpch.PART_ROWS_PER_SF * sf)
    li = tpch.gen_lineitem(rows, nparts, seed=42, device="cuda")
    part = tpch.gen_part(nparts, seed=43, device="cuda")
    hli = {k: v.cpu().pin_memory() for k, v in li.items()}
    hpart = {k: v.cpu().pin_memory() for k, v in part.items()}
    del li, part
    torch.cuda.empty_cache()
    rv1, rv14, pt = bench.host_tables(hli, hpart, rows)
    p1, p14 = bench.plans(rv1, rv14, pt)
    b1, b14 = split_rowvector(rv1, 1 << 26), split_rowvector(rv14, 1 << 26)
    for rep in range(2):
        cache = UploadCache()
        t0 = time.perf_counter()
        t1 = Task(p1)
        t1.set_upload_cache(cache)
        for b in b1:
            t1.add_input(0, b)
        ta = time.perf_counter()
        t1.run()
        tb = time.perf_counter()
        s1 = t1.stats()
        t1.close()
        t14 = Task(p14)
        t14.set_upload_cache(cache)
        for b in b14:
            t14.add_input(0, b)
        t14.add_input(1, pt)
        tc = time.perf_counter()
        t14.run()
        td = time.perf_counter()
        s14 = t14.stats()
        t14.close()
        cache.close()
        h2d = s1.get("task.h2dBytes", 0) + s14.get("task.h2dBytes", 0)
        out[f"e2e_rep{rep}"] = {"rows": rows, "total_s": td - t0, "q1_add_input_s": ta - t0, "q1_run_s": tb - ta, "q14_add_input_s": tc - tb, "q14_run_s": td - tc,
                               "h2d_bytes": h2d, "GBps": h2d / (td - t0) / 1e9,
                               "q1_wall_ms": {k: round(v / 1e6, 2) for k, v in s1.items() if k.endswith("WallNanos") and v > 1e6},
                               "q14_wall_ms": {k: round(v / 1e6, 2) for k, v in s14.items() if k.endswith("WallNanos") and v > 1e6}}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
