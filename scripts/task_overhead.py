"""Where the time of one query through the operator-level C ABI goes on the host side:
create / add_input / run / result / close, per phase, next to the operators' own wall times
(VB2_SYNC_TIMING=1 attributes kernel time to the operator that launched it)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from velox_b200 import tpch
from velox_b200.task import Task, _result


def phases(plan, inputs, iters):
    acc = {"create": 0.0, "add_input": 0.0, "run": 0.0, "result": 0.0, "stats": 0.0, "close": 0.0}
    for it in range(iters + 3):
        t0 = time.perf_counter()
        t = Task(plan)
        t1 = time.perf_counter()
        for sid, cols in inputs:
            t.add_input(sid, cols)
        t2 = time.perf_counter()
        import ctypes as C
        err = C.create_string_buffer(2048)
        rc = t.L.vb2_task_run(t.h, err, 2048)
        assert rc == 0, err.value
        t3 = time.perf_counter()
        out = _result(t.L, t.h, plan.names)
        t4 = time.perf_counter()
        st = t.stats()
        t5 = time.perf_counter()
        t.close()
        t6 = time.perf_counter()
        if it >= 3:
            for k, v in zip(acc, (t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t6 - t5)):
                acc[k] += v
    return {k: round(v / iters * 1e3, 4) for k, v in acc.items()}, {k: round(v / 1e6, 3) for k, v in st.items() if k.endswith("WallNanos") and v > 5e3}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=10)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    rows = int(tpch.LINEITEM_ROWS_PER_SF * a.sf)
    nparts = int(tpch.PART_ROWS_PER_SF * a.sf)
    li = tpch.gen_lineitem(rows, nparts, device="cuda")
    part = tpch.gen_part(nparts, device="cuda")
    small = {k: v[:1000].cpu() for k, v in li.items()}
    rv1, rv14, pt = bench.host_tables(small, {k: v[:1000].cpu() for k, v in part.items()}, 1000)
    q1, q14 = bench.plans(rv1, rv14, pt)
    c1, c14, cp = bench.device_inputs(li, part)
    torch.cuda.synchronize()
    out = {"sf": a.sf, "rows": rows, "sync_timing": os.environ.get("VB2_SYNC_TIMING", "0")}
    for name, plan, inputs in (("q1", q1, [(0, c1)]), ("q14", q14, [(0, c14), (1, cp)])):
        ph, ops = phases(plan, inputs, a.iters)
        out[name] = {"phases_ms": ph, "total_ms": round(sum(ph.values()), 4), "operators_ms": ops}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
