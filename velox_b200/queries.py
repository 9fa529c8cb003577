"""TPC-H Q1 / Q6 / Q14 on device-resident columns, at the kernel-level C ABI (one process per GPU).

Plans: velox/exec/tests/utils/TpchQueryBuilder.cpp:203-256 (Q1), :756-788 (Q6), :1639-1702 (Q14).
The same kernels are what the operator layer launches for these plans (tests assert
`b200.fusedBatches`); this module adds the multi-GPU composition of SURVEY.md §8e:

  Q1 / Q6   rows sharded by range; per-GPU partial aggregates; one tiny NCCL all-reduce
            (partialAggregation -> localPartition({}) -> finalAggregation).
  Q14       both sides hash-partitioned by VectorHasher-hash(key) % world
            (velox/exec/HashPartitionFunction.cpp:113-116): each GPU filters + projects its
            lineitem shard into (l_partkey, revenue) with the fused scan-compact kernel, scatters
            rows and its part shard into per-peer segments, ONE grouped ncclSend/ncclRecv
            all-to-all per column, then local build + fused probe + CASE + sums, final 2-value
            all-reduce.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from . import tpch
from .kernels import (DeviceColumn, FusedScanAgg, FusedScanCompact, LikeOnAlphabet, column_minmax, flat_device, gather,
                      hash_columns, join_build_array, join_slot_flags, key_range_check, normalize_keys, partition_ids,
                      partition_scatter_order, partition_segments)
from .vector import BIGINT

Q14_SCAN_SIG = "F:between(i0,pi0,pi1);C:l1|multiply(f2,minus(pf0,f3))"
Q14_PROBE_SIG = "F:true;P:f1|switch(joinflag,f1,pf0);J:l0"


class Q1:
    """sum/avg/count over (l_returnflag, l_linestatus): one fused kernel, 44 B/row."""
    NGROUPS = len(tpch.RETURNFLAG_DICT) * len(tpch.LINESTATUS_DICT)

    def __init__(self, comm=None):
        self.f = FusedScanAgg(tpch.Q1_SIG, ngroups=self.NGROUPS)
        self.comm = comm

    def launch(self, li, rows):
        self.f.reset()
        self.f.add_batch([li["l_shipdate"], li["l_quantity"], li["l_extendedprice"], li["l_discount"], li["l_tax"]], rows,
                         pf=[1.0, 1.0, 1.0], pi=[tpch.Q1_SHIPDATE_LT], keys=[li["l_returnflag"], li["l_linestatus"]],
                         key_min=[0, 0], key_mult=[len(tpch.LINESTATUS_DICT), 1])

    def merge(self):
        if self.comm is not None:  # partial -> final across GPUs: <= 6 groups x 6 values
            self.comm.all_reduce_(self.f.sums)
            self.comm.all_reduce_(self.f.counts)

    def result(self):
        sums = self.f.sums.cpu().numpy().reshape(self.NGROUPS, 5)
        counts = self.f.counts.cpu().numpy()
        out = {}
        for g in range(self.NGROUPS):
            c = int(counts[g])
            if c == 0:
                continue
            s = sums[g]
            key = (tpch.RETURNFLAG_DICT[g // len(tpch.LINESTATUS_DICT)], tpch.LINESTATUS_DICT[g % len(tpch.LINESTATUS_DICT)])
            out[key] = (s[0], s[1], s[2], s[3], s[0] / c, s[1] / c, s[4] / c, c)
        return out


class Q6:
    def __init__(self, comm=None):
        self.f = FusedScanAgg(tpch.Q6_SIG)
        self.comm = comm

    def launch(self, li, rows):
        self.f.reset()
        self.f.add_batch([li["l_shipdate"], li["l_discount"], li["l_quantity"], li["l_extendedprice"]], rows,
                         pf=[0.05, 0.07, 24.0], pi=[tpch.Q6_SHIP_LO, tpch.Q6_SHIP_HI])

    def merge(self):
        if self.comm is not None:
            self.comm.all_reduce_(self.f.sums)
            self.comm.all_reduce_(self.f.counts)

    def result(self):
        return self.f.sums.item() if self.f.counts.item() else None


class Q14:
    """100 * sum(case when p_type like 'PROMO%' then rev else 0 end) / sum(rev) over lineitem |x| part."""

    def __init__(self, comm=None, compact_capacity: Optional[int] = None, overlap: Optional[bool] = None):
        self.comm = comm
        # build side of the planned exchange on a second stream, concurrent with the lineitem scan
        self.overlap = (os.environ.get("VB2_Q14_OVERLAP", "1") == "1") if overlap is None else overlap
        self.side = None
        self.like = LikeOnAlphabet(tpch.PTYPE_DICT, "PROMO%")
        if comm is None or comm.world == 1:
            self.probe = FusedScanAgg(tpch.Q14_SIG)
        else:
            self.probe = FusedScanAgg(Q14_PROBE_SIG)
            self.compact_capacity = compact_capacity
            self.scan = None
            self.plan = None       # statistics-derived sizes of the planned (sync-free) execution
            self.overflow = None
            self.planned_runs = 0
            self._last = None

    # ---- build side -----------------------------------------------------------------------------
    def _build(self, partkey: torch.Tensor, ptype: torch.Tensor):
        """Array-mode table over the build keys -> one byte per key slot (0 miss, 1 match, 2 match & PROMO)."""
        n = partkey.numel()
        if n == 0:
            return torch.zeros(1, dtype=torch.uint8, device="cuda"), 0
        col = flat_device(BIGINT, partkey)
        lo, hi, _ = column_minmax(col)  # key range decides the layout (VectorHasher range mode)
        rng = hi - lo + 2
        keys, valid = normalize_keys([col], [lo], [1], ranges=[rng], nulls_invalid=True)
        head, _next, _flags = join_build_array(keys, valid, rng)
        flags = self.like.run()
        return join_slot_flags(head, ptype, flags), lo - 1  # slot = key - (lo - 1)

    # ---- single GPU -----------------------------------------------------------------------------
    def launch(self, li, part, rows):
        if self.comm is not None and self.comm.world > 1:
            return self._launch_partitioned(li, part, rows)
        slot_flags, join_min = self._build(part["p_partkey"], part["p_type"])
        self.probe.reset()
        self.probe.add_batch([li["l_shipdate"], li["l_partkey"], li["l_extendedprice"], li["l_discount"]], rows,
                             pf=[1.0, 1.0, 0.0], pi=[tpch.Q14_SHIP_LO, tpch.Q14_SHIP_HI],
                             join={"slot_flags": slot_flags, "min": join_min})

    # ---- hash-partitioned across GPUs -----------------------------------------------------------------
    # Two executions of the same exchange plan:
    #   * planning run (first launch, and again after an overflow): sizes are discovered on the way —
    #     per-peer counts travel to the host before each all-to-all, the build key range comes from a
    #     device min/max. Its statistics (largest segment of either side, key range, all-reduced so
    #     every rank plans the same sizes) are kept.
    #   * planned run (every later launch): fixed-capacity segments sized from those statistics with
    #     head-room, tails padded with a sentinel key that misses every probe / build. No count
    #     exchange, no host synchronisation: the whole query is one asynchronous launch sequence
    #     whose only host read is the final result. Any statistic that no longer holds (a segment
    #     or the scan output overflows, a key outside the planned range) raises a device flag that
    #     result() checks; the query then reruns as a planning run.
    HEADROOM = 1.25

    def _exchange(self, key: torch.Tensor, payload: torch.Tensor):
        """Partitions rows by VectorHasher-hash(key) % world and exchanges both columns
        (launches on the current stream; one host synchronisation for the counts)."""
        w = self.comm.world
        h = hash_columns([flat_device(BIGINT, key)])
        ids = partition_ids(h, w)
        counts, order = partition_scatter_order(ids, w)
        sk, sp = gather(key, order), gather(payload, order)
        send_counts, recv_counts = self.comm.exchange_counts_dev(counts)
        k, p = self.comm.all_to_all_columns([sk, sp], send_counts, recv_counts)
        return k, p, max(send_counts)

    def _launch_partitioned(self, li, part_shard, rows):
        self._last = (li, part_shard, rows)
        if self.plan is not None:
            return self._launch_planned(li, part_shard, rows)
        if self.scan is None:
            self.scan = FusedScanCompact(Q14_SCAN_SIG, self.compact_capacity or max(1 << 20, rows // 16))
        pad = lambda v: max(64, (int(v * self.HEADROOM) + 63) // 64 * 64)
        while True:  # the planning run discovers the scan output size too
            self.scan.err.zero_()
            self.scan.run([li["l_shipdate"], li["l_partkey"], li["l_extendedprice"], li["l_discount"]], rows,
                          pf=[1.0], pi=[tpch.Q14_SHIP_LO, tpch.Q14_SHIP_HI])
            n = int(self.scan.count.item())
            if n <= self.scan.capacity:
                break
            self.scan = FusedScanCompact(Q14_SCAN_SIG, pad(n))
        n, (lk, rev) = self.scan.result([torch.int64, torch.float64])
        # NCCL calls are issued in the same order on every rank (part exchange, then lineitem exchange).
        pk, pt, part_seg = self._exchange(part_shard["p_partkey"], part_shard["p_type"])
        slot_flags, join_min = self._build(pk, pt)
        rk, rrev, li_seg = self._exchange(lk, rev)
        self.probe.reset()
        m = rk.numel()
        if m:
            self.probe.add_batch([rk, rrev], m, pf=[0.0], join={"slot_flags": slot_flags, "min": join_min})
        # statistics for the planned runs, identical on every rank
        lo, hi, cnt = column_minmax(flat_device(BIGINT, pk)) if pk.numel() else (0, -1, 0)
        big = 1 << 62
        mx = torch.tensor([li_seg, part_seg, n, hi if cnt else -big, -(lo if cnt else big)], dtype=torch.int64, device="cuda")
        self.comm.all_reduce_max_(mx)
        li_seg, part_seg, n_max, hi, neg_lo = mx.tolist()
        lo = -neg_lo
        if hi < lo:
            return  # no build rows anywhere: nothing to plan
        self.plan = {"li_seg": pad(li_seg), "part_seg": pad(part_seg), "lo": lo, "hi": hi}
        if n_max * self.HEADROOM > self.scan.capacity:
            self.scan = FusedScanCompact(Q14_SCAN_SIG, pad(n_max))

    def _launch_planned(self, li, part_shard, rows):
        w, pl = self.comm.world, self.plan
        if self.overflow is None:
            self.overflow = torch.zeros(2, dtype=torch.int64, device="cuda")
        self.flag = self.scan.err  # one device flag for every broken assumption: scan output, segments, key range
        main = torch.cuda.current_stream()
        self.flag.zero_()
        if self.overlap:
            if self.side is None:
                self.side = torch.cuda.Stream()
            ready = torch.cuda.Event()
            ready.record(main)  # inputs and the cleared flag are complete here
        self.scan.run([li["l_shipdate"], li["l_partkey"], li["l_extendedprice"], li["l_discount"]], rows,
                      pf=[1.0], pi=[tpch.Q14_SHIP_LO, tpch.Q14_SHIP_HI])

        def build_side():
            # part rows to the owners of their keys, then the array-mode table over the planned key range
            seg = [pl["part_seg"]] * w
            pk_seg, (pt_seg,), _ = partition_segments(part_shard["p_partkey"], [part_shard["p_type"]], part_shard["p_partkey"].numel(), None, w,
                                                      pl["part_seg"], self.flag)
            pk, pt = self.comm.all_to_all_columns([pk_seg, pt_seg], seg, seg)
            key_range_check(pk, pl["lo"], pl["hi"], self.flag)
            rng = pl["hi"] - pl["lo"] + 2
            keys, valid = normalize_keys([flat_device(BIGINT, pk)], [pl["lo"]], [1], ranges=[rng], nulls_invalid=True)  # sentinel -> invalid
            head, _next, _f = join_build_array(keys, valid, rng)
            return join_slot_flags(head, pt, self.like.run())

        if self.overlap:
            # the build side runs on a second stream while the scan streams HBM; NCCL calls keep the
            # same order on every rank (part exchange, then lineitem exchange)
            self.side.wait_event(ready)
            with torch.cuda.stream(self.side):
                slot_flags = build_side()
                built = torch.cuda.Event()
                built.record()
        else:
            slot_flags = build_side()
        # probe side: the compacted (l_partkey, revenue) rows, row count read on the device
        lk, rev = self.scan.outs[0].view(torch.int64), self.scan.outs[1].view(torch.float64)
        seg = [pl["li_seg"]] * w
        lk_seg, (rev_seg,), _ = partition_segments(lk, [rev], self.scan.capacity, self.scan.count, w, pl["li_seg"], self.flag)
        rk, rrev = self.comm.all_to_all_columns([lk_seg, rev_seg], seg, seg)
        if self.overlap:
            main.wait_event(built)
            slot_flags.record_stream(main)
        self.probe.reset()
        self.probe.add_batch([rk, rrev], rk.numel(), pf=[0.0], join={"slot_flags": slot_flags, "min": pl["lo"] - 1})
        self.planned_runs += 1

    def merge(self):
        if self.comm is not None and self.comm.world > 1:
            self.comm.all_reduce_(self.probe.sums)
            self.comm.all_reduce_(self.probe.counts)
            if self.plan is not None and self.overflow is not None:
                # a broken planning assumption anywhere invalidates the run everywhere
                self.overflow.copy_(self.flag)
                self.comm.all_reduce_(self.overflow)

    def result(self):
        if self.comm is not None and self.comm.world > 1 and self.plan is not None and self.overflow is not None \
                and int(self.overflow.sum().item()) != 0:
            # the data outgrew the plan: run again discovering sizes, which also re-plans
            self.plan = None
            self.overflow = None
            self.scan = None
            self.compact_capacity = None
            li, part_shard, rows = self._last
            self._launch_partitioned(li, part_shard, rows)
            self.merge()
        total, promo = self.probe.sums.cpu().tolist()
        return (100.0 * promo / total) if total else None
