"""Repeats the post-exchange probe pipeline on a large input and reports run-to-run spread (development aid)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from velox_b200 import tpch
from velox_b200.kernels import FusedScanAgg
from velox_b200.queries import Q14, Q14_PROBE_SIG

rows, nparts = int(os.environ.get("ROWS", 6_000_000)), 50_000
g = torch.Generator(device="cuda"); g.manual_seed(1)
k = torch.randint(1, nparts + 1, (rows,), generator=g, device="cuda", dtype=torch.int64)
v = torch.rand(rows, generator=g, device="cuda", dtype=torch.float64) * 1e5
if os.environ.get("MODE") == "lineitem":
    li = tpch.gen_lineitem(rows, nparts, seed=42, device="cuda")
    k = li["l_partkey"].clone()
    v = (li["l_extendedprice"] * (1.0 - li["l_discount"])).clone()
    del li
part = tpch.gen_part(nparts, seed=43, device="cuda")
q = Q14()
slot_flags, join_min = q._build(part["p_partkey"], part["p_type"])
probe = FusedScanAgg(Q14_PROBE_SIG)
vals = []
for rep in range(int(os.environ.get("REPS", 20))):
    probe.reset()
    probe.add_batch([k, v], rows, pf=[0.0], join={"slot_flags": slot_flags, "min": join_min})
    vals.append(probe.sums[0].item())
want = float(v.sum())
errs = [abs(x - want) / want for x in vals]
print(json.dumps({"mode": os.environ.get("MODE"), "rows": rows, "dbg": os.environ.get("VB2_FUSED_DBG"), "stages": os.environ.get("VB2_FUSED_STAGES"), "distinct_results": len(set(vals)), "max_rel_err": max(errs), "bad_runs": sum(e > 1e-12 for e in errs)}))
