"""Stage-by-stage check of the partitioned-Q14 building blocks on ONE GPU against torch restatements
(development aid): scan-compact, hash partitioning (order-based and fixed segments), post-exchange probe."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from velox_b200 import tpch
from velox_b200.kernels import (SENTINEL_KEY, FusedScanAgg, FusedScanCompact, flat_device, gather, hash_columns, partition_ids,
                                partition_scatter_order, partition_segments)
from velox_b200.queries import Q14, Q14_PROBE_SIG, Q14_SCAN_SIG
from velox_b200.vector import BIGINT


def main():
    rows, nparts, W = int(os.environ.get("ROWS", 3_000_000)), 50_000, 2
    out = {}
    li = tpch.gen_lineitem(rows, nparts, seed=42, device="cuda")
    if os.environ.get("ALLPASS", "1") == "1":
        li["l_shipdate"] = torch.full_like(li["l_shipdate"], tpch.Q14_SHIP_LO)
    part = tpch.gen_part(nparts, seed=43, device="cuda")
    q = Q14()
    slot_flags, join_min = q._build(part["p_partkey"], part["p_type"])

    def probe_check(tag, k, v, reps=4):
        probe = FusedScanAgg(Q14_PROBE_SIG)
        vals = []
        for rep in range(reps):
            probe.reset()
            probe.add_batch([k, v], k.numel(), pf=[0.0], join={"slot_flags": slot_flags, "min": join_min})
            vals.append(probe.sums[0].item())
        live = k != SENTINEL_KEY
        want = float(v[live].sum())
        out["probe@" + tag] = {"distinct": len(set(vals)), "max_rel_err": max(abs(x - want) / want for x in vals)}

    k0 = li["l_partkey"].clone()
    v0 = (li["l_extendedprice"] * (1.0 - li["l_discount"])).clone()
    probe_check("start", k0, v0)
    # 1. scan-compact
    scan = FusedScanCompact(Q14_SCAN_SIG, rows + 1024)
    for rep in range(3):
        scan.run([li["l_shipdate"], li["l_partkey"], li["l_extendedprice"], li["l_discount"]], rows, pf=[1.0], pi=[tpch.Q14_SHIP_LO, tpch.Q14_SHIP_HI])
        n, (lk, rev) = scan.result([torch.int64, torch.float64])
        m = (li["l_shipdate"] >= tpch.Q14_SHIP_LO) & (li["l_shipdate"] <= tpch.Q14_SHIP_HI)
        wk = li["l_partkey"][m]
        wr = (li["l_extendedprice"] * (1.0 - li["l_discount"]))[m]
        o1, o2 = torch.argsort(lk * 1), torch.argsort(wk)
        same_keys = n == wk.numel() and bool((lk[o1] == wk[o2]).all())
        # (key, rev) pairs as multisets: compare sums per key
        a = torch.zeros(nparts + 1, dtype=torch.float64, device="cuda").index_add_(0, lk, rev)
        b = torch.zeros(nparts + 1, dtype=torch.float64, device="cuda").index_add_(0, wk, wr)
        out[f"compact_rep{rep}"] = {"n": n, "want": wk.numel(), "keys_equal": same_keys, "max_rel_key_sum_diff": float(((a - b).abs() / b.abs().clamp_min(1)).max())}
    probe_check("after_compact_fresh", k0, v0)
    probe_check("after_compact_views", lk, rev)
    lk, rev = lk.clone(), rev.clone()
    probe_check("after_compact_clones", lk, rev)
    # 2. order-based partitioning
    h = hash_columns([flat_device(BIGINT, lk)])
    ids = partition_ids(h, W)
    counts, order = partition_scatter_order(ids, W)
    sk, sp = gather(lk, order), gather(rev, order)
    c = counts.tolist()
    ok = sum(c) == n
    off = 0
    for p in range(W):
        seg_ids = ids[order[off:off + c[p]].long()]
        ok = ok and bool((seg_ids == p).all())
        off += c[p]
    ok = ok and bool((torch.sort(order.long()).values == torch.arange(n, device="cuda")).all())
    out["partition_order"] = {"ok": ok, "counts": c}
    probe_check("after_partition_order", lk, rev)
    # 3. fixed segments
    flag = torch.zeros(2, dtype=torch.int32, device="cuda")
    segcap = (max(c) * 5 // 4 + 63) // 64 * 64
    cnt_dev = torch.tensor([n], dtype=torch.int64, device="cuda")
    big_k = torch.cat([lk, torch.full((5000,), 7, dtype=torch.int64, device="cuda")])  # rows beyond *rows_dev must be ignored
    big_v = torch.cat([rev, torch.zeros(5000, dtype=torch.float64, device="cuda")])
    sk2, (sv2,), cnt2 = partition_segments(big_k, [big_v], big_k.numel(), cnt_dev, W, segcap, flag)
    ok = cnt2.tolist() == c and int(flag[0].item()) == 0
    off = 0
    for p in range(W):
        seg_k = sk2[p * segcap:(p + 1) * segcap]
        ok = ok and bool((seg_k[:c[p]] == sk[off:off + c[p]]).all()) and bool((seg_k[c[p]:] == SENTINEL_KEY).all())
        ok = ok and bool((sv2[p * segcap:p * segcap + c[p]] == sp[off:off + c[p]]).all())
        off += c[p]
    out["partition_segments"] = {"ok": ok, "counts": cnt2.tolist(), "segcap": segcap}
    probe_check("after_segments", lk, rev)
    probe_check("after_segments_fresh", k0, v0)
    probe_check("segments", sk2, sv2)
    # 4. post-exchange probe over all rows (one "rank" owning every key)
    if os.environ.get("LATE_BUILD") == "1":
        q = Q14()
        slot_flags, join_min = q._build(part["p_partkey"], part["p_type"])
    promo_codes = torch.tensor([1 if s.startswith("PROMO") else 0 for s in tpch.PTYPE_DICT], device="cuda")
    is_promo = promo_codes[part["p_type"].long()].bool()
    flag_by_key = torch.zeros(nparts + 2, dtype=torch.bool, device="cuda")
    flag_by_key[part["p_partkey"]] = is_promo
    for name, (k, v) in {"dense": (lk, rev), "segments": (sk2, sv2)}.items():
        probe = FusedScanAgg(Q14_PROBE_SIG)
        vals = []
        for rep in range(3):
            probe.reset()
            probe.add_batch([k, v], k.numel(), pf=[0.0], join={"slot_flags": slot_flags, "min": join_min})
            vals.append(probe.sums.cpu().tolist() + [int(probe.counts.item())])
        live = k != SENTINEL_KEY
        want_total = float(v[live].sum())
        want_promo = float(v[live][flag_by_key[k[live]]].sum())
        out[f"probe_{name}"] = {"got": vals, "want": [want_total, want_promo, int(live.sum())]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
