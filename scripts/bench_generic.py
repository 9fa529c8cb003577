"""Operator path on device-resident input: Q1 / Q6 / Q14 through the Task API with the fused
pipelines on and off (off = B200FilterProject VM kernels + generic aggregation / join kernels)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import bench
from velox_b200 import tpch
from velox_b200.kernels import DeviceColumn, flat_device
from velox_b200.task import Task
from velox_b200.vector import BIGINT, DOUBLE, INTEGER, VARCHAR, DICTIONARY, dictionary_vector


def dict_dev(codes, alphabet):
    host = dictionary_vector(VARCHAR, torch.zeros(1, dtype=torch.int32).numpy(), alphabet)
    d = DeviceColumn.from_host(host)
    d.indices, d.size = codes, codes.numel()
    return d


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sf", type=float, default=10)
    a = ap.parse_args()
    rows = int(tpch.LINEITEM_ROWS_PER_SF * a.sf)
    nparts = int(tpch.PART_ROWS_PER_SF * a.sf)
    li = tpch.gen_lineitem(rows, nparts, device="cuda")
    part = tpch.gen_part(nparts, device="cuda")
    small = {k: v[:1000].cpu() for k, v in li.items()}
    rv1, rv14, pt = bench.host_tables(small, {k: v[:1000].cpu() for k, v in part.items()}, 1000)
    q1, q14 = bench.plans(rv1, rv14, pt)
    c1 = [dict_dev(li["l_returnflag"], tpch.RETURNFLAG_DICT), dict_dev(li["l_linestatus"], tpch.LINESTATUS_DICT)] + \
         [flat_device(DOUBLE, li[c]) for c in ("l_quantity", "l_extendedprice", "l_discount", "l_tax")] + [flat_device(INTEGER, li["l_shipdate"])]
    c14 = [flat_device(BIGINT, li["l_partkey"]), flat_device(DOUBLE, li["l_extendedprice"]), flat_device(DOUBLE, li["l_discount"]), flat_device(INTEGER, li["l_shipdate"])]
    cp = [flat_device(BIGINT, part["p_partkey"]), dict_dev(part["p_type"], tpch.PTYPE_DICT)]
    torch.cuda.synchronize()
    out = {"sf": a.sf, "rows": rows}
    for name, plan, inputs in (("q1", q1, [(0, c1)]), ("q14", q14, [(0, c14), (1, cp)])):
        for label, cfg in (("fused", {}), ("generic", {"b200.fused_pipelines": "false"})):
            ts = []
            for it in range(4):
                t = Task(plan, cfg)
                for sid, cols in inputs:
                    t.add_input(sid, cols)
                t0 = time.perf_counter()
                r = t.run()
                ts.append(time.perf_counter() - t0)
                st = t.stats()
                t.close()
            ms = sorted(ts[1:])[1] * 1e3
            out[f"{name}_{label}"] = {"ms": ms, "rows_per_s": rows / ms * 1e3, "result_rows": r.size,
                                       "wall_ms": {k: round(v / 1e6, 3) for k, v in st.items() if k.endswith("WallNanos") and v > 2e4}}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
