#!/usr/bin/env python
"""bench.py — TPC-H Q1 & Q14 at SF100 on in-HBM columns (BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W            our arm (N>1: under torchrun)
  python bench.py --impl reference --steps K --warmup W    CPU arm: the oracle (restatement of the
                                                           reference's velox/exec CPU operators)

A step = one pass of the hot path over the resident workload: Q1 (filter + 2-key group-by, 8
aggregates) followed by Q14 (filter + project, hash join with part, CASE, 2 sums) over the same
SF100 lineitem columns. rows/s counts the lineitem rows scanned by the two queries. Scaling is
strong: SF100 in total, row-sharded over the GPUs; Q1 merges per-GPU partials with one tiny
all-reduce, Q14 hash-partitions both join sides with one NCCL all-to-all per column.

`value`   kernels + collectives on data already in HBM (CUDA events, max over ranks).
`e2e`     the same two plans through the operator-level C ABI (vb2_task_*: Task -> Driver ->
          B200 operators) with HOST (pinned) input columns: host->device copies of every input
          column and the device->host read of the result are inside the timed region.
`roofline` the dominant kernel (Q1's fused scan-filter-project-aggregate): algorithmic bytes
          (44 B/row) / CUDA-event time of that launch, vs the measured HBM peak.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "rows/sec TPC-H Q1 & Q14 SF100"
UNIT = "rows/s"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled through NVML every ~5 ms while the GPU is under load."""

    def __init__(self, index: int):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.samples, self.max_mhz = index, threading.Event(), [], None
        # NVML queries serialise with CUDA driver calls: polling every few ms stretches the launch
        # sequence of a query measurably (0.3-0.4 ms per step at a 5 ms period), so the period is 25 ms
        self.period = float(os.environ.get("VB2_BENCH_SAMPLE_MS", "25")) / 1e3

    def run(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            while not self.stop_flag.is_set():
                self.samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM),
                                     pynvml.nvmlDeviceGetCurrentClocksEventReasons(h), 0))
                self.stop_flag.wait(self.period)
        except Exception as e:  # pragma: no cover
            self.error = str(e)

    def summary(self):
        import pynvml
        sm = sorted(s[0] for s in self.samples)
        flags = {"hw_slowdown": pynvml.nvmlClocksEventReasonHwSlowdown, "hw_thermal_slowdown": pynvml.nvmlClocksEventReasonHwThermalSlowdown,
                 "sw_thermal_slowdown": pynvml.nvmlClocksEventReasonSwThermalSlowdown, "sw_power_cap": pynvml.nvmlClocksEventReasonSwPowerCap}
        reasons = sorted(n for n, bit in flags.items() if any(s[1] & bit for s in self.samples))
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": self.max_mhz, "reasons": reasons, "samples": len(sm)}


# -------------------------------------------------------------------------------------------------
# CPU arm: the oracle on the host cores
# -------------------------------------------------------------------------------------------------
def host_tables(li, part, rows):
    from velox_b200 import tpch
    from velox_b200.vector import BIGINT, DOUBLE, INTEGER, VARCHAR, dictionary_vector, flat_vector, row_vector
    h = {k: v[:rows].numpy() if hasattr(v, "numpy") else v[:rows] for k, v in li.items()}
    q1_names = ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate"]

    def col(n):
        if n == "l_returnflag":
            return dictionary_vector(VARCHAR, h[n], tpch.RETURNFLAG_DICT)
        if n == "l_linestatus":
            return dictionary_vector(VARCHAR, h[n], tpch.LINESTATUS_DICT)
        if n == "l_shipdate":
            return flat_vector(INTEGER, h[n])
        if n == "l_partkey":
            return flat_vector(BIGINT, h[n])
        return flat_vector(DOUBLE, h[n])

    rv1 = row_vector(q1_names, [col(n) for n in q1_names])
    q14_names = ["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"]
    rv14 = row_vector(q14_names, [col(n) for n in q14_names])
    p = {k: (v.numpy() if hasattr(v, "numpy") else v) for k, v in part.items()}
    pt = row_vector(["p_partkey", "p_type"], [flat_vector(BIGINT, p["p_partkey"]), dictionary_vector(VARCHAR, p["p_type"], tpch.PTYPE_DICT)])
    return rv1, rv14, pt


def plans(rv1, rv14, pt):
    from velox_b200.plan import PlanBuilder
    q1 = (PlanBuilder().values(rv1.names, rv1.types)
          .filter("l_shipdate < '1998-09-03'::DATE")
          .project(["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice",
                    "l_extendedprice * (1.0 - l_discount) AS l_sum_disc_price",
                    "l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS l_sum_charge", "l_discount"])
          .partialAggregation(["l_returnflag", "l_linestatus"],
                              ["sum(l_quantity)", "sum(l_extendedprice)", "sum(l_sum_disc_price)", "sum(l_sum_charge)",
                               "avg(l_quantity)", "avg(l_extendedprice)", "avg(l_discount)", "count(0)"])
          .localPartition([]).finalAggregation().planNode())
    build = PlanBuilder().values(pt.names, pt.types, source=1)
    q14 = (PlanBuilder().values(rv14.names, rv14.types, source=0)
           .filter("l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE")
           .project(["l_extendedprice * (1.0 - l_discount) as part_revenue", "l_shipdate", "l_partkey"])
           .hashJoin(["l_partkey"], ["p_partkey"], build, "", ["part_revenue", "p_type"])
           .project(["(CASE WHEN (p_type LIKE 'PROMO%') THEN part_revenue ELSE 0.0 END) as filter_revenue", "part_revenue"])
           .partialAggregation([], ["sum(part_revenue) as total_revenue", "sum(filter_revenue) as total_promo_revenue"])
           .localPartition([]).finalAggregation()
           .project(["100.00 * total_promo_revenue/total_revenue as promo_revenue"]).planNode())
    return q1, q14


def plans_distributed(rv1, rv14, pt):
    """The same queries as multi-fragment plans, one process per GPU: partial aggregates are gathered
    by a PartitionedOutput -> Exchange pair in front of the final aggregation; Q14 hash-partitions both
    join sides by the join key (HashPartitionFunction) before the local build + probe
    (velox/exec/tests/MultiFragmentTest.cpp builds such plans for the reference)."""
    from velox_b200.plan import PlanBuilder
    q1 = (PlanBuilder().values(rv1.names, rv1.types)
          .filter("l_shipdate < '1998-09-03'::DATE")
          .project(["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice",
                    "l_extendedprice * (1.0 - l_discount) AS l_sum_disc_price",
                    "l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS l_sum_charge", "l_discount"])
          .partialAggregation(["l_returnflag", "l_linestatus"],
                              ["sum(l_quantity)", "sum(l_extendedprice)", "sum(l_sum_disc_price)", "sum(l_sum_charge)",
                               "avg(l_quantity)", "avg(l_extendedprice)", "avg(l_discount)", "count(0)"])
          .gatherExchange().finalAggregation().planNode())
    build = PlanBuilder().values(pt.names, pt.types, source=1).partitionedOutput(["p_partkey"])
    q14 = (PlanBuilder().values(rv14.names, rv14.types, source=0)
           .filter("l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE")
           .project(["l_extendedprice * (1.0 - l_discount) as part_revenue", "l_partkey"])
           .partitionedOutput(["l_partkey"])
           .hashJoin(["l_partkey"], ["p_partkey"], build, "", ["part_revenue", "p_type"])
           .project(["(CASE WHEN (p_type LIKE 'PROMO%') THEN part_revenue ELSE 0.0 END) as filter_revenue", "part_revenue"])
           .partialAggregation([], ["sum(part_revenue) as total_revenue", "sum(filter_revenue) as total_promo_revenue"])
           .gatherExchange().finalAggregation()
           .project(["100.00 * total_promo_revenue/total_revenue as promo_revenue"]).planNode())
    return q1, q14


def cpu_sample(sf_rows: int, nparts: int, sample_rows: int, seed=42):
    import torch
    from velox_b200 import tpch
    li = tpch.gen_lineitem(sample_rows, nparts, seed=seed, device="cpu")
    part = tpch.gen_part(nparts, seed=43, device="cpu")
    return li, part


def run_cpu(rv1, rv14, pt, threads: int, batch_rows=10000):
    """One step of the CPU arm: Q1 then Q14 through the oracle's batch-at-a-time drivers."""
    from oracle import pyoracle
    q1, q14 = plans(rv1, rv14, pt)
    t0 = time.perf_counter()
    r1 = pyoracle.run_plan(q1, [rv1], threads=threads, batch_rows=batch_rows)
    r14 = pyoracle.run_plan(q14, [rv14, pt], threads=threads, batch_rows=batch_rows)
    return time.perf_counter() - t0, r1, r14


def cpu_build_seconds(rv1, rv14, pt, threads: int):
    """Q14 over a one-row probe side: the cost of the hash build over the whole part table, which a
    lineitem sample pays in full. Used to scale the sample's time to the full workload."""
    from oracle import pyoracle
    from velox_b200.task import split_rowvector
    _, q14 = plans(rv1, rv14, pt)
    tiny = split_rowvector(rv14, 1, max_batches=1)[0]  # (splitting the whole sample into one-row batches took minutes)
    best = float("inf")
    for _ in range(2):  # the first call also pays one-time costs: keep the faster one
        t0 = time.perf_counter()
        pyoracle.run_plan(q14, [tiny, pt], threads=threads, batch_rows=10000)
        best = min(best, time.perf_counter() - t0)
    return best


def cpu_rows_per_s(sec_sample: float, sec_build: float, sample: int, rows_total: int):
    """rows/s of the full workload from a lineitem sample: scan-proportional time scales with the
    rows, the part-table build does not (the sample joins against the whole part table)."""
    sec_build = min(sec_build, 0.9 * sec_sample)
    full = (sec_sample - sec_build) * (rows_total / sample) + sec_build
    return 2 * rows_total / full


def reference_arm(args):
    """--impl reference: the reference's CPU path for the same metric. The real velox/exec cannot be
    built in this environment (folly/fmt/xsimd/DuckDB absent), so this arm times the
    oracle — the CPU restatement of those operators — on all host threads over a bounded sample."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from velox_b200 import tpch
    threads = os.cpu_count() or 1
    sample = int(args.cpu_sample_rows)
    nparts = int(tpch.PART_ROWS_PER_SF * args.sf)
    li, part = cpu_sample(0, nparts, sample)
    rv1, rv14, pt = host_tables(li, part, sample)
    for _ in range(args.warmup):
        run_cpu(rv1, rv14, pt, threads)
    times = []
    for _ in range(args.steps):
        t, _, _ = run_cpu(rv1, rv14, pt, threads)
        times.append(t)
    sec = sum(times) / len(times)
    rows_total = int(tpch.LINEITEM_ROWS_PER_SF * args.sf)
    sec_build = cpu_build_seconds(rv1, rv14, pt, threads)
    value = cpu_rows_per_s(sec, sec_build, sample, rows_total)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TPC-H Q1+Q14, {sample} lineitem rows of the SF{args.sf:g} columns per step (bounded sample), part {nparts} rows",
                   "reference_build": "velox/exec not buildable here; oracle = CPU restatement of its operators"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
                         "sample": f"{sample} of {rows_total} lineitem rows x (Q1 + Q14) against the whole part table, 10K-row batches, {threads} driver "
                                   f"threads: {sec:.2f} s per step of which {sec_build:.2f} s is the part-table build; scaled to the full workload as "
                                   f"(step - build) x {rows_total / sample:.2f} + build"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# -------------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------------
def device_inputs(li, part):
    """Column batches already resident in HBM, as the operator-level C ABI takes them (VB2_DEVICE)."""
    import torch
    from velox_b200 import tpch
    from velox_b200.kernels import DeviceColumn, flat_device
    from velox_b200.vector import BIGINT, DOUBLE, INTEGER, VARCHAR, dictionary_vector

    def dict_dev(codes, alphabet):
        host = dictionary_vector(VARCHAR, torch.zeros(1, dtype=torch.int32).numpy(), alphabet)
        d = DeviceColumn.from_host(host)
        d.indices, d.size = codes, codes.numel()
        return d

    c1 = [dict_dev(li["l_returnflag"], tpch.RETURNFLAG_DICT), dict_dev(li["l_linestatus"], tpch.LINESTATUS_DICT)] + \
         [flat_device(DOUBLE, li[c]) for c in ("l_quantity", "l_extendedprice", "l_discount", "l_tax")] + [flat_device(INTEGER, li["l_shipdate"])]
    c14 = [flat_device(BIGINT, li["l_partkey"]), flat_device(DOUBLE, li["l_extendedprice"]), flat_device(DOUBLE, li["l_discount"]),
           flat_device(INTEGER, li["l_shipdate"])]
    cp = [flat_device(BIGINT, part["p_partkey"]), dict_dev(part["p_type"], tpch.PTYPE_DICT)]
    return c1, c14, cp


def slice_inputs(cols, n):
    """First n rows of device column batches (views, no copy)."""
    from velox_b200.kernels import DeviceColumn
    out = []
    for c in cols:
        d = DeviceColumn(c.type, c.encoding, n, values=c.values if c.encoding else c.values[:n], nulls=None,
                         indices=None if c.indices is None else c.indices[:n], dict_size=c.dict_size, dict_nulls=c.dict_nulls, aux=c.aux)
        out.append(d)
    return out


def run_task(plan, inputs, config=None, comm=None):
    """One query through the operator-level C ABI: vb2_task_create / add_input(VB2_DEVICE) / run / result."""
    from velox_b200.task import Task
    t = Task(plan, config)
    try:
        if comm is not None:
            t.set_comm(comm)
        for sid, cols in inputs:
            t.add_input(sid, cols)
        out = t.run()
        return out, t.stats()
    finally:
        t.close()


def q1_rows(rv):
    """(returnflag, linestatus) -> (4 sums, 3 avgs, count) from a Q1 result RowVector."""
    return {(r[0], r[1]): tuple(r[2:]) for r in rv.rows()}


def parity(got1, got14, want1, want14):
    """GPU result vs CPU oracle over the same rows: integer / key columns exact, FP
    columns by maximum relative error."""
    g, w = q1_rows(got1), q1_rows(want1)
    exact = set(g) == set(w)
    worst = 0.0
    for k in w:
        if k not in g:
            continue
        exact = exact and int(g[k][7]) == int(w[k][7])
        for a, b in zip(g[k][:7], w[k][:7]):
            worst = max(worst, abs(a - b) / max(abs(b), 1e-300))
    a, b = got14.rows()[0][0], want14.rows()[0][0]
    worst = max(worst, abs(a - b) / max(abs(b), 1e-300))
    return {"exact_columns_ok": bool(exact), "fp_max_rel_err": worst, "fp_tolerance": 1e-9,
            "ok": bool(exact and worst <= 1e-9)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--sf", type=float, default=100.0)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--concurrent-queries", action="store_true",
                    help="submit Q1 and Q14 as two concurrent tasks (vb2_tasks_run) instead of one after the other. Off by default: on 2 GPUs a query's "
                         "exchange rendezvous hides behind the other query's scan (4.63 vs 5.44 ms per step), but on 4 GPUs the two tasks' exchanges "
                         "interleave differently on every rank and the step degrades to 30 ms (profiles/r02_bench_n4.json) against 3.9 ms one after the other")
    ap.add_argument("--serial-queries", action="store_true", help="run Q1, then Q14 (the default)")
    ap.add_argument("--cpu-sample-rows", type=float, default=60_000_000)
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        return reference_arm(args)

    import ctypes
    import torch
    import torch.distributed as dist
    from velox_b200 import tpch
    from velox_b200._lib import lib
    from velox_b200.queries import Q1, Q6, Q14

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.serial_queries:
        args.concurrent_queries = False
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a GPU (there is no CPU fallback)"
    torch.cuda.set_device(local_rank)
    comm = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        from velox_b200.comm import Comm
        comm = Comm()
    L = lib()
    L.vb2k_kernel_launches.restype = ctypes.c_int64

    rows_total = int(tpch.LINEITEM_ROWS_PER_SF * args.sf) + (2 if args.sf == 100 else 0)  # SF100 = 600 037 902
    nparts = int(tpch.PART_ROWS_PER_SF * args.sf)
    r0, r1 = rows_total * rank // world, rows_total * (rank + 1) // world
    rows = r1 - r0
    p0, p1 = nparts * rank // world, nparts * (rank + 1) // world
    li = tpch.gen_lineitem(rows, nparts, seed=42 + rank, device="cuda")
    part_all = tpch.gen_part(nparts, seed=43, device="cuda")
    part = {k: v[p0:p1].contiguous() for k, v in part_all.items()} if world > 1 else part_all
    torch.cuda.synchronize()

    # plans (the reference's TpchQueryBuilder shapes) and device-resident inputs for the operator API
    small = {k: v[:1000].cpu() for k, v in li.items()}
    rv1s, rv14s, pts = host_tables(small, {k: v[:1000].cpu() for k, v in part_all.items()}, 1000)
    plan1, plan14 = plans(rv1s, rv14s, pts)
    c1, c14, cp = device_inputs(li, part)

    state = {"q1_kernel_ns": 0, "q1_kernel_rows": 0, "q1_runs": 0}

    # A step runs Q1, then Q14. With --concurrent-queries it submits them as two concurrent tasks (one host thread each,
    # each with its own streams and — on N > 1 — its own communicator): see the flag's help for what was measured.
    comm14 = None
    if world > 1:
        from velox_b200.comm import Comm
        comm14 = Comm()  # exchanges of concurrent tasks must not share epochs
        dplan1, dplan14 = plans_distributed(rv1s, rv14s, pts)
        pl1, pl14 = dplan1, dplan14
    else:
        pl1, pl14 = plan1, plan14

    from velox_b200.task import Task, run_tasks

    def make(plan, inputs, c):
        t = Task(plan)
        if c is not None:
            t.set_comm(c)
        for sid, cols in inputs:
            t.add_input(sid, cols)
        return t

    def step(record=False):
        t1, t14 = make(pl1, [(0, c1)], comm), make(pl14, [(0, c14), (1, cp)], comm14)
        try:
            if (not args.concurrent_queries):
                out1, out14 = t1.run(), t14.run()
            else:
                out1, out14 = run_tasks([t1, t14])  # vb2_tasks_run: both tasks at once, one library thread each
            if record:
                # only the fused-scan timing is read every step (the full stats text is parsed once, on the last step)
                s1 = t1.stats(only="b200.fusedScan")
                state["q1_kernel_ns"] += sum(v for k, v in s1.items() if k.endswith("b200.fusedScanNanos"))
                state["q1_kernel_rows"] += sum(v for k, v in s1.items() if k.endswith("b200.fusedScanRows"))
                state["q1_runs"] += 1
                if record == "last":
                    state["stats1"], state["stats14"] = t1.stats(), t14.stats()
        finally:
            t1.close()
            t14.close()
        return out1, out14

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()  # samples through warm-up, the timed region and the per-query loops (all under load)
    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = L.vb2k_kernel_launches()
    start, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t_wall = time.perf_counter()
    start.record()
    for i in range(args.steps):
        out1, out14 = step(record="last" if i == args.steps - 1 else True)
    end.record()
    barrier()
    wall_ms = (time.perf_counter() - t_wall) * 1e3 / args.steps
    launches = L.vb2k_kernel_launches() - launches0
    ms = start.elapsed_time(end) / args.steps
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    value = 2 * rows_total / (ms / 1e3)

    # dominant kernel: Q1's fused scan + its finalize, timed by the operator with CUDA events on the
    # stream it launches on (b200.fusedScanNanos), inside the timed steps above
    q1_ms = state["q1_kernel_ns"] / max(1, state["q1_runs"]) / 1e6
    peak, peak_src = peaks()
    achieved = rows * tpch.Q1_BYTES_PER_ROW / (q1_ms / 1e3) / 1e9 if q1_ms else 0.0
    traffic, traffic_src = None, None
    for name in ("r02_q1_dram.json", "r01_q1_dram.json"):  # newest ncu --set full capture of this kernel
        prof = os.path.join(ROOT, "profiles", name)
        if os.path.exists(prof):
            with open(prof) as f:
                pj = json.load(f)
            traffic = pj["dram_bytes_per_row"] * rows
            traffic_src = f"profiles/{name}: dram__bytes_read.sum + dram__bytes_write.sum of one ncu --set full capture of this kernel, per row x rows of this launch"
            break
    roofline = {"bound": "hbm", "kernel": "fused_scan_agg_tma_kernel<Q1>", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "kernel_ms": q1_ms, "algorithmic_bytes": rows * tpch.Q1_BYTES_PER_ROW,
                "timing": "CUDA events recorded by B200HashAggregation on its launch stream around the fused launch (stat b200.fusedScanNanos)"}

    # ---- side by side: operator level (Task API) vs kernel level (vb2k_* called directly) ---------------
    def timed_wall(fn, iters=10):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        return (time.perf_counter() - t0) * 1e3 / iters

    def timed_events(fn, iters=40):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(iters):
            fn()
        b.record()
        b.synchronize()
        return a.elapsed_time(b) / iters

    breakdown = {}
    q1k, q14k, q6k = Q1(comm), Q14(comm), Q6(comm)
    kernel_level = {"q1": lambda: (q1k.launch(li, rows), q1k.merge()), "q14": lambda: (q14k.launch(li, part, rows), q14k.merge()),
                    "q6": lambda: (q6k.launch(li, rows), q6k.merge())}
    bprs = {"q1": tpch.Q1_BYTES_PER_ROW, "q14": tpch.Q14_BYTES_PER_ROW, "q6": tpch.Q6_BYTES_PER_ROW}
    for name, fn in kernel_level.items():
        m = timed_events(fn)
        mm = torch.tensor([m], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(mm, op=dist.ReduceOp.MAX)
        m = float(mm.item())
        breakdown[name] = {"kernel_level_ms": m, "rows_per_s_kernel_level": rows_total / (m / 1e3),
                           "algorithmic_GBps_per_gpu": rows * bprs[name] / (m / 1e3) / 1e9, "frac_of_hbm_peak": rows * bprs[name] / (m / 1e3) / 1e9 / peak}
    # each query alone through the operator API (on N > 1 every rank takes part: the exchanges are collective)
    breakdown["q1"]["operator_level_ms"] = timed_wall(lambda: run_task(pl1, [(0, c1)], comm=comm))
    breakdown["q14"]["operator_level_ms"] = timed_wall(lambda: run_task(pl14, [(0, c14), (1, cp)], comm=comm14))
    sampler.stop_flag.set()
    sampler.join()

    results = {"q1": {f"{k[0]}{k[1]}": int(v[7]) for k, v in q1_rows(out1).items()}, "q14_promo_revenue": out14.rows()[0][0] if out14.size else None,
               "q6_revenue": q6k.result()}
    # kernel-level and operator-level answers must agree (rank 0 holds the gathered operator-level result)
    k1 = q1k.result()
    if rank == 0:
        assert {f"{k[0]}{k[1]}": int(v[7]) for k, v in k1.items()} == results["q1"], "operator-level and kernel-level Q1 counts differ"
        k14 = q14k.result()
        assert abs(results["q14_promo_revenue"] - k14) <= 1e-9 * abs(k14), ("operator-level and kernel-level Q14 differ", results["q14_promo_revenue"], k14)

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "wall_ms_per_step": wall_ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": f"TPC-H Q1 + Q14 over SF{args.sf:g} in-HBM lineitem ({rows_total} rows) and part ({nparts} rows)",
                   "path": "operator-level C ABI: vb2_task_create / add_input(VB2_DEVICE) / run / result per query (Task -> Driver -> B200 operators); "
                           "the timed step contains plan parsing, operator setup, every kernel, and the result rows on the host",
                   "queries_per_step": "Q1 and Q14 submitted as two concurrent tasks" if not (not args.concurrent_queries) else "Q1, then Q14",
                   "rows_per_gpu": rows, "parallelism": "1 GPU" if world == 1 else f"row-sharded x{world}; Q14 hash-partitioned, NCCL all-to-all",
                   "l2": "inputs (26-31 GB per pass) far exceed the 126 MB L2; no flush needed", "value_counts": "2 x lineitem rows per step"},
        "roofline": roofline, "queries": breakdown, "results": results, "gpu_launches": int(launches),
        "clocks": sampler.summary(),
    }
    if comm is not None:
        ex1, ex14 = comm.exchanges(), comm14.exchanges()
        line["exchange"] = {"transport": "peer memory over NVLink (CUDA IPC heaps, exchange_p2p.cu)" if comm.peer_memory else "NCCL grouped send/recv",
                            "exchanges": {k: ex1[k] + ex14[k] for k in ex1}}
    if "stats1" in state:
        line["operator_wall_ms"] = {q: {k: round(v / 1e6, 3) for k, v in state[s].items() if k.endswith("WallNanos") and v > 1e4}
                                    for q, s in (("q1", "stats1"), ("q14", "stats14")) if s in state}

    # ---- e2e: operator-level C ABI with host buffers ----------------------------------------------------
    if not args.skip_e2e:
        try:
            line["e2e"] = e2e(args, li, part_all, rows, nparts, world, rank, rows_total)
        except Exception as ex:  # keep the device-timed line even if the end-to-end leg cannot run here
            line["e2e"] = {"value": None, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0, "error": f"{type(ex).__name__}: {ex}"[:300]}
    # ---- CPU baseline + parity (rank 0, N = 1) -----------------------------------------------------------
    if world == 1 and not args.skip_cpu:
        threads = os.cpu_count() or 1
        sample = int(min(args.cpu_sample_rows, rows))
        hli = {k: v[:sample].cpu() for k, v in li.items()}
        rv1, rv14, pt = host_tables(hli, {k: v.cpu() for k, v in part_all.items()}, sample)
        run_cpu(rv1, rv14, pt, threads)  # warm-up
        sec, r1c, r14c = run_cpu(rv1, rv14, pt, threads)
        sec_build = cpu_build_seconds(rv1, rv14, pt, threads)
        line["cpu_baseline"] = {"value": cpu_rows_per_s(sec, sec_build, sample, rows_total), "unit": UNIT, "cores": threads, "kind": "port",
                                "sample": f"first {sample} of {rows_total} lineitem rows x (Q1 + Q14) against the whole part table, 10K-row batches, "
                                          f"{threads} driver threads: {sec:.2f} s of which {sec_build:.2f} s is the part-table build; scaled to the full "
                                          f"workload as (step - build) x {rows_total / sample:.2f} + build"}
        # the same rows through the B200 operators, compared with what the CPU arm just computed
        g1, _ = run_task(plan1, [(0, slice_inputs(c1, sample))])
        g14, _ = run_task(plan14, [(0, slice_inputs(c14, sample)), (1, cp)])
        line["parity"] = dict(parity(g1, g14, r1c, r14c), rows=sample, oracle="CPU oracle (restatement of velox/exec; operator results of the reference are "
                              "pinned by DuckDB at its test time, which is absent here)")
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def gpu_local_cpus(index: int):
    """CPUs of the NUMA node the GPU hangs off (sysfs local_cpulist of its PCI function), or None."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = (bus.decode() if isinstance(bus, bytes) else bus).lower()[-12:]
        text = open(f"/sys/bus/pci/devices/{bus}/local_cpulist").read().strip()
        cpus = set()
        for part in text.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        return cpus or None
    except Exception:
        return None


def h2d_probe_gbps(nbytes=2 << 30):
    import torch
    host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    host.fill_(1)
    dev = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    dev.copy_(host, non_blocking=True)
    torch.cuda.synchronize()
    return nbytes / (time.perf_counter() - t0) / 1e9


def e2e(args, li, part_all, rows, nparts, world, rank, rows_total):
    """Q1 + Q14 through vb2_task_* with pinned HOST columns: every step copies all input columns to
    the device (B200FromHost) and reads the result rows back (B200ToHost)."""
    import torch
    import torch.distributed as dist
    from velox_b200.task import Task, split_rowvector

    # Pinned host columns live on the GPU's own NUMA node (page-locked memory is placed by the allocating
    # thread's affinity; a cross-socket source halves the PCIe rate on two-socket hosts). The affinity is
    # restored afterwards: the CPU arm uses every core.
    before = os.sched_getaffinity(0)
    local = gpu_local_cpus(int(os.environ.get("LOCAL_RANK", "0")))
    if local:
        os.sched_setaffinity(0, local)
    try:
        hli = {k: v.cpu().pin_memory() for k, v in li.items()}
        hpart = {k: v.cpu().pin_memory() for k, v in part_all.items()}
        probe = h2d_probe_gbps()
    finally:
        os.sched_setaffinity(0, before)
    rv1, rv14, pt = host_tables(hli, hpart, rows)
    p1, p14 = plans(rv1, rv14, pt)
    batch = 1 << 26
    b1, b14 = split_rowvector(rv1, batch), split_rowvector(rv14, batch)
    h2d = sum(c.values.nbytes if c.encoding == 0 else c.indices.nbytes for c in rv1.columns) + \
        sum(c.values.nbytes if c.encoding == 0 else c.indices.nbytes for c in rv14.columns) + \
        sum(c.values.nbytes if c.encoding == 0 else c.indices.nbytes for c in pt.columns)

    from velox_b200.task import UploadCache
    state = {}

    def one():
        # Both queries read the same host lineitem batches: an upload cache that lives for this step
        # lets Q14 reuse the device copies of the columns Q1 already brought over (l_extendedprice,
        # l_discount, l_shipdate). Everything is uploaded again in the next step.
        cache = UploadCache()
        t1 = Task(p1)
        t1.set_upload_cache(cache)
        for b in b1:
            t1.add_input(0, b)
        out1 = t1.run()
        s1 = t1.stats()
        t1.close()
        t14 = Task(p14)
        t14.set_upload_cache(cache)
        for b in b14:
            t14.add_input(0, b)
        t14.add_input(1, pt)
        out14 = t14.run()
        s14 = t14.stats()
        t14.close()
        cache.close()
        state["h2d"] = s1.get("task.h2dBytes", 0) + s14.get("task.h2dBytes", 0)
        return out1, out14

    one()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(args.e2e_steps):
        out1, out14 = one()
    sec = (time.perf_counter() - t0) / args.e2e_steps
    t = torch.tensor([sec], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    sec = float(t.item())
    d2h = sum(len(r) * 8 for r in out1.rows()) + 8
    note = "per-rank Task over its row shard; cross-rank merge of the (tiny) results not included" if world > 1 else "full plan through one Task"
    copied = int(state.get("h2d", 0)) or int(h2d)  # bytes the tasks actually copied (shared columns once per step)
    return {"value": 2 * rows_total / sec, "unit": UNIT, "h2d_bytes_per_step": copied * world, "d2h_bytes_per_step": int(d2h) * world,
            "ms_per_step": sec * 1e3, "steps": args.e2e_steps, "h2d_GBps_effective": copied / sec / 1e9,
            "h2d_GBps_link_probe": probe, "link_probe": "one 2 GiB pinned host -> device copy on this box, timed alone (what the PCIe link delivers)",
            "path": "vb2_task_create/add_input(HOST)/run; pinned host columns, 64M-row batches; per-step upload cache shared by the two tasks",
            "h2d_bytes_without_sharing": int(h2d) * world, "note": note}


if __name__ == "__main__":
    main()
