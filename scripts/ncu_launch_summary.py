"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time and launch
count per kernel name (cold-cache, serialised launches: shares are meaningful, absolute times are
upper bounds)."""
import csv
import json
import re
import sys
from collections import defaultdict


def main(path, skip=0):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        if r.get("Metric Name") != "gpu__time_duration.sum":
            continue
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        ns = v * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)
        rows.append((int(r["ID"]), r["Kernel Name"], ns))
    rows = rows[skip:]
    agg = defaultdict(lambda: [0.0, 0])
    for _, name, ns in rows:
        short = re.sub(r"<.*", "", name)
        agg[short][0] += ns
        agg[short][1] += 1
    total = sum(v[0] for v in agg.values())
    out = [{"kernel": k, "launches": v[1], "total_us": round(v[0] / 1e3, 1), "avg_us": round(v[0] / v[1] / 1e3, 2), "share": round(v[0] / total, 4)}
           for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])]
    print(json.dumps({"launches": len(rows), "total_us": round(total / 1e3, 1), "kernels": out}, indent=1))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
