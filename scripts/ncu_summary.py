"""Extracts the metrics the roofline discussion needs from an .ncu-rep (run where ncu is installed)."""
import csv
import json
import subprocess
import sys

WANT = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "launch__shared_mem_per_block_dynamic", "smsp__inst_executed.sum",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "lts__t_sector_hit_rate.pct",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "smsp__average_warp_latency_issue_stalled_long_scoreboard.pct"]


def main(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    res = []
    for vals in rows[2:]:
        d = {"kernel": vals[hdr.index("Kernel Name")]}
        for h, u, v in zip(hdr, units, vals):
            if h in WANT:
                d[h] = f"{v} {u}".strip()
        res.append(d)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main(sys.argv[1])
