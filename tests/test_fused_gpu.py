"""Fused scan->filter->project->aggregate kernels (sm_100a) against the CPU oracle.

Mirrors the plans of velox/exec/tests/utils/TpchQueryBuilder.cpp (Q1 :203-256, Q6 :756-788,
Q14 :1639-1702). Counts are compared bit-exactly, SUM(double) within relative 1e-12 (the kernel
sums in a fixed tree order, the reference sequentially in input order)."""
import os

import numpy as np
import pytest
import torch

from oracle import pyoracle
from velox_b200 import tpch
from velox_b200.plan import PlanBuilder
from velox_b200.vector import (BIGINT, DOUBLE, INTEGER, VARCHAR, dictionary_vector, flat_vector, row_vector)

pytestmark = pytest.mark.gpu


def rel_tol(n):
    """Stated FP tolerance for SUM/AVG(double): the reference adds sequentially in input order,
    which carries up to (n-1)*2^-53 relative error on same-sign data; the kernel's fixed tree
    order carries ~log2(n)*2^-53. Results agree within max(1e-12, n*2^-53) relative."""
    return max(1e-12, n * 2.0 ** -53)


def _host_lineitem(n, nparts=2000, seed=7):
    t = tpch.gen_lineitem(n, nparts, seed=seed, device="cpu")
    return {k: v.numpy() for k, v in t.items()}


def _lineitem_rowvector(h, names):
    cols = []
    for nme in names:
        if nme == "l_returnflag":
            cols.append(dictionary_vector(VARCHAR, h[nme], tpch.RETURNFLAG_DICT))
        elif nme == "l_linestatus":
            cols.append(dictionary_vector(VARCHAR, h[nme], tpch.LINESTATUS_DICT))
        elif nme == "l_shipdate":
            cols.append(flat_vector(INTEGER, h[nme]))
        elif nme == "l_partkey":
            cols.append(flat_vector(BIGINT, h[nme]))
        else:
            cols.append(flat_vector(DOUBLE, h[nme]))
    return row_vector(names, cols)


def _dev(h, name):
    return torch.from_numpy(h[name]).cuda()


@pytest.mark.parametrize("n", [0, 1, 2, 3, 255, 4096, 100_003, 1 << 20])
def test_q6_matches_oracle(n):
    from velox_b200.kernels import FusedScanAgg
    h = _host_lineitem(max(n, 1))
    h = {k: v[:n] for k, v in h.items()}
    names = ["l_shipdate", "l_extendedprice", "l_quantity", "l_discount"]
    rv = _lineitem_rowvector(h, names)
    plan = (PlanBuilder().values(rv.names, rv.types)
            .filter("l_shipdate between '1994-01-01'::DATE and '1994-12-31'::DATE and "
                    "l_discount between 0.05 and 0.07 and l_quantity < 24.0")
            .project(["l_extendedprice * l_discount"])
            .partialAggregation([], ["sum(p0)"]).localPartition([]).finalAggregation().planNode())
    want = pyoracle.run_plan(plan, [rv]).rows()[0][0]
    f = FusedScanAgg(tpch.Q6_SIG)
    if n:
        f.add_batch([_dev(h, "l_shipdate"), _dev(h, "l_discount"), _dev(h, "l_quantity"), _dev(h, "l_extendedprice")],
                    n, pf=[0.05, 0.07, 24.0], pi=[tpch.Q6_SHIP_LO, tpch.Q6_SHIP_HI])
    got, cnt = f.sums.cpu().item(), f.counts.cpu().item()
    m = ((h["l_shipdate"] >= tpch.Q6_SHIP_LO) & (h["l_shipdate"] <= tpch.Q6_SHIP_HI) & (h["l_discount"] >= 0.05)
         & (h["l_discount"] <= 0.07) & (h["l_quantity"] < 24.0))
    assert cnt == int(m.sum())  # filter mask: bit exact
    if cnt == 0:
        assert want is None and got == 0.0
    else:
        assert abs(got - want) <= rel_tol(n) * abs(want)
        # and within 1e-13 of the correctly rounded sum of the same per-row products
        import math
        exact = math.fsum((h["l_extendedprice"][m] * h["l_discount"][m]).tolist())
        assert abs(got - exact) <= 1e-13 * abs(exact)


def test_q6_unaligned_and_multibatch():
    """Batches starting at odd row offsets take the scalar-load variant; accumulators persist."""
    from velox_b200.kernels import FusedScanAgg
    n = 50_001
    h = _host_lineitem(n)
    d = {k: _dev(h, k) for k in ["l_shipdate", "l_discount", "l_quantity", "l_extendedprice"]}
    f = FusedScanAgg(tpch.Q6_SIG)
    cuts = [0, 1, 4098, 33_333, n]
    for a, b in zip(cuts[:-1], cuts[1:]):
        f.add_batch([d["l_shipdate"][a:b], d["l_discount"][a:b], d["l_quantity"][a:b], d["l_extendedprice"][a:b]],
                    b - a, pf=[0.05, 0.07, 24.0], pi=[tpch.Q6_SHIP_LO, tpch.Q6_SHIP_HI])
    m = ((h["l_shipdate"] >= tpch.Q6_SHIP_LO) & (h["l_shipdate"] <= tpch.Q6_SHIP_HI) & (h["l_discount"] >= 0.05)
         & (h["l_discount"] <= 0.07) & (h["l_quantity"] < 24.0))
    want = float(np.sum(h["l_extendedprice"][m] * h["l_discount"][m]))
    assert f.counts.item() == int(m.sum())
    assert abs(f.sums.item() - want) <= rel_tol(n) * abs(want)


def test_q6_nan_ordering():
    """NaN is the largest value in comparisons (type/FloatingPointUtil.h:52-98): `qty < 24` is
    false for NaN, `disc between 0.05 and 0.07` is false for NaN."""
    from velox_b200.kernels import FusedScanAgg
    ship = torch.full((8,), tpch.Q6_SHIP_LO, dtype=torch.int32, device="cuda")
    disc = torch.tensor([0.06, float("nan"), 0.06, 0.05, 0.07, 0.08, 0.06, 0.06], dtype=torch.float64, device="cuda")
    qty = torch.tensor([1.0, 1.0, float("nan"), 23.0, 24.0, 1.0, float("inf"), -float("inf")], dtype=torch.float64, device="cuda")
    ep = torch.arange(1, 9, dtype=torch.float64, device="cuda")
    f = FusedScanAgg(tpch.Q6_SIG)
    f.add_batch([ship, disc, qty, ep], 8, pf=[0.05, 0.07, 24.0], pi=[tpch.Q6_SHIP_LO, tpch.Q6_SHIP_HI])
    # rows kept: 0 (1*0.06), 3 (4*0.05), 7 (8*0.06)
    assert f.counts.item() == 3
    assert f.sums.item() == pytest.approx(1 * 0.06 + 4 * 0.05 + 8 * 0.06, rel=1e-15)


@pytest.mark.parametrize("n", [1, 2, 1000, 65_537, 1 << 20])
def test_q1_matches_oracle(n):
    from velox_b200.kernels import FusedScanAgg
    h = _host_lineitem(n, seed=11)
    names = ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate"]
    rv = _lineitem_rowvector(h, names)
    plan = (PlanBuilder().values(rv.names, rv.types)
            .filter("l_shipdate < '1998-09-03'::DATE")
            .project(["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice",
                      "l_extendedprice * (1.0 - l_discount) AS l_sum_disc_price",
                      "l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS l_sum_charge", "l_discount"])
            .partialAggregation(["l_returnflag", "l_linestatus"],
                                ["sum(l_quantity)", "sum(l_extendedprice)", "sum(l_sum_disc_price)", "sum(l_sum_charge)",
                                 "avg(l_quantity)", "avg(l_extendedprice)", "avg(l_discount)", "count(0)"])
            .localPartition([]).finalAggregation().planNode())
    want = {(r[0], r[1]): r[2:] for r in pyoracle.run_plan(plan, [rv]).rows()}
    f = FusedScanAgg(tpch.Q1_SIG, ngroups=6)
    f.add_batch([_dev(h, "l_shipdate"), _dev(h, "l_quantity"), _dev(h, "l_extendedprice"), _dev(h, "l_discount"), _dev(h, "l_tax")],
                n, pf=[1.0, 1.0, 1.0], pi=[tpch.Q1_SHIPDATE_LT],
                keys=[_dev(h, "l_returnflag"), _dev(h, "l_linestatus")], key_min=[0, 0], key_mult=[2, 1])
    sums = f.sums.cpu().numpy().reshape(6, 5)
    counts = f.counts.cpu().numpy()
    got = {}
    for g in range(6):
        if counts[g] == 0:
            continue
        key = (tpch.RETURNFLAG_DICT[g // 2], tpch.LINESTATUS_DICT[g % 2])
        s = sums[g]
        c = int(counts[g])
        got[key] = (s[0], s[1], s[2], s[3], s[0] / c, s[1] / c, s[4] / c, c)
    assert set(got) == set(want)
    for key, w in want.items():
        g = got[key]
        assert g[7] == w[7]  # count: bit exact
        for a, b in zip(g[:7], w[:7]):
            assert abs(a - b) <= rel_tol(n) * abs(b), (key, a, b)


def test_q14_matches_oracle():
    from velox_b200.kernels import FusedScanAgg, join_slot_flags
    n, nparts = 300_000, 5000
    h = _host_lineitem(n, nparts=nparts, seed=3)
    part = {k: v.numpy() for k, v in tpch.gen_part(nparts, seed=5).items()}
    li = _lineitem_rowvector(h, ["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"])
    pt = row_vector(["p_partkey", "p_type"], [flat_vector(BIGINT, part["p_partkey"]),
                                             dictionary_vector(VARCHAR, part["p_type"], tpch.PTYPE_DICT)])
    build = PlanBuilder().values(pt.names, pt.types, source=1)
    plan = (PlanBuilder().values(li.names, li.types, source=0)
            .filter("l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE")
            .project(["l_extendedprice * (1.0 - l_discount) as part_revenue", "l_shipdate", "l_partkey"])
            .hashJoin(["l_partkey"], ["p_partkey"], build, "", ["part_revenue", "p_type"])
            .project(["(CASE WHEN (p_type LIKE 'PROMO%') THEN part_revenue ELSE 0.0 END) as filter_revenue", "part_revenue"])
            .partialAggregation([], ["sum(part_revenue) as total_revenue", "sum(filter_revenue) as total_promo_revenue"])
            .localPartition([]).finalAggregation()
            .project(["100.00 * total_promo_revenue/total_revenue as promo_revenue"]).planNode())
    want = pyoracle.run_plan(plan, [li, pt]).rows()[0][0]
    # build side: dense array table key -> row + 1; build-side predicate per dictionary entry
    head = torch.zeros(nparts, dtype=torch.int32, device="cuda")
    pk = torch.from_numpy(part["p_partkey"]).cuda()
    head[pk - 1] = torch.arange(1, nparts + 1, dtype=torch.int32, device="cuda")
    flag = torch.tensor([1 if s.startswith("PROMO") else 0 for s in tpch.PTYPE_DICT], dtype=torch.uint8, device="cuda")
    f = FusedScanAgg(tpch.Q14_SIG)
    f.add_batch([_dev(h, "l_shipdate"), _dev(h, "l_partkey"), _dev(h, "l_extendedprice"), _dev(h, "l_discount")], n,
                pf=[1.0, 1.0, 0.0], pi=[tpch.Q14_SHIP_LO, tpch.Q14_SHIP_HI],
                join={"slot_flags": join_slot_flags(head, torch.from_numpy(part["p_type"]).cuda(), flag), "min": 1})
    total, promo = f.sums.cpu().tolist()
    got = 100.00 * promo / total
    assert abs(got - want) <= 1e-12 * abs(want)
    m = (h["l_shipdate"] >= tpch.Q14_SHIP_LO) & (h["l_shipdate"] <= tpch.Q14_SHIP_HI)
    assert f.counts.item() == int(m.sum())


def test_hash_and_partition_match_oracle():
    """VectorHasher::hash / HashPartitionFunction (exec/VectorHasher.cpp:62-126,
    exec/HashPartitionFunction.cpp:75-118): bit exact, incl. nulls, NaN, -0.0, dictionary,
    constant and string keys of every length class of bits::hashBytes."""
    from velox_b200.kernels import DeviceColumn, hash_columns, partition_ids, partition_scatter_order
    from velox_b200.vector import constant_vector
    rng = np.random.default_rng(5)
    n = 20_000
    i64 = rng.integers(-2**62, 2**62, n)
    i64[:4] = [0, -1, 2**63 - 1, -2**63]
    i32 = rng.integers(-2**31, 2**31 - 1, n).astype(np.int32)
    f64 = rng.standard_normal(n)
    f64[:6] = [0.0, -0.0, np.nan, -np.nan, np.inf, -np.inf]
    nulls = rng.random(n) < 0.1
    strs = ["".join(chr(97 + (i * 7 + j) % 26) for j in range(i % 41)) for i in range(n)]
    dict_idx = rng.integers(0, 150, n).astype(np.int32)
    host = [
        flat_vector(BIGINT, i64, nulls),
        flat_vector(INTEGER, i32),
        flat_vector(DOUBLE, f64),
        flat_vector(VARCHAR, strs),
        dictionary_vector(VARCHAR, dict_idx, tpch.PTYPE_DICT),
        constant_vector(BIGINT, 42, n),
        constant_vector(DOUBLE, None, n),
    ]
    dev = [DeviceColumn.from_host(c) for c in host]
    for pick in ([0], [1], [2], [3], [4], [0, 1, 2], [3, 4, 5, 6], list(range(7))):
        want = pyoracle.hash_columns([host[i] for i in pick])
        got = hash_columns([dev[i] for i in pick]).cpu().numpy().view(np.uint64)
        assert np.array_equal(got, want), pick
    want_h = pyoracle.hash_columns([host[0], host[3]])
    got_h = hash_columns([dev[0], dev[3]])
    for p in (1, 2, 7, 8, 64):
        want = pyoracle.partition([host[0], host[3]], p)
        ids = partition_ids(got_h, p)
        assert np.array_equal(ids.cpu().numpy().view(np.uint32), want)
        counts, order = partition_scatter_order(ids, p)
        order = order.cpu().numpy()
        counts = counts.cpu().numpy()
        assert np.array_equal(counts, np.bincount(want, minlength=p))
        # stable grouping by partition
        assert np.array_equal(order, np.argsort(want, kind="stable").astype(np.int32))


@pytest.mark.parametrize("p", [2, 5, 64])
def test_partition_order_large(p):
    """Stable partition order over 10 M rows (2 442 blocks of 4 096 rows: the offsets scan gives every thread a run of
    blocks) against numpy's stable argsort; ids given and ids computed on the fly from one BIGINT key."""
    import torch
    from velox_b200.kernels import partition_scatter_order
    rng = np.random.default_rng(p)
    n = 10_000_019
    ids = rng.integers(0, p, n).astype(np.int32)
    ids[: n // 3] = 1 % p  # a long run of one partition
    counts, order = partition_scatter_order(torch.from_numpy(ids).cuda(), p)
    assert np.array_equal(counts.cpu().numpy(), np.bincount(ids, minlength=p))
    assert np.array_equal(order.cpu().numpy(), np.argsort(ids, kind="stable").astype(np.int32))


@pytest.mark.parametrize("card", [(2, 2), (3, 2), (5, 1), (8, 5), (50, 1)])
@pytest.mark.parametrize("n", [1000, 300_001])
def test_fused_groupby_variants(card, n):
    """Register accumulators (<= 4 groups) and shared-memory accumulators (more groups), 32- and
    64-bit keys, against numpy: counts bit exact, sums within the stated tolerance."""
    from velox_b200.kernels import FusedScanAgg
    rng = np.random.default_rng(card[0] * 100 + card[1] + n)
    k0 = rng.integers(10, 10 + card[0], n).astype(np.int32)
    k1 = rng.integers(-3, -3 + card[1], n).astype(np.int32)
    x = np.round(rng.uniform(-1000, 1000, n), 2)
    G = card[0] * card[1]
    for dtype in (np.int32, np.int64):
        f = FusedScanAgg("F:true;P:f0", ngroups=G)
        f.add_batch([torch.from_numpy(x).cuda()], n, keys=[torch.from_numpy(k0.astype(dtype)).cuda(), torch.from_numpy(k1.astype(dtype)).cuda()],
                    key_min=[10, -3], key_mult=[card[1], 1])
        gid = (k0 - 10) * card[1] + (k1 + 3)
        want_cnt = np.bincount(gid, minlength=G)
        want_sum = np.bincount(gid, weights=x, minlength=G)
        assert np.array_equal(f.counts.cpu().numpy(), want_cnt)
        got = f.sums.cpu().numpy()
        assert np.all(np.abs(got - want_sum) <= 1e-9 + rel_tol(n) * np.abs(want_sum))


@pytest.mark.parametrize("rows", [6_000_000 + 321])
def test_tma_stage_reuse_is_exact_and_deterministic(rows):
    """Every block of the TMA-staged kernels cycles through its shared-memory stages many times at
    this size. A stage may only be handed back to the producer once every value loaded from it is
    in a register; a violation shows up as stale tiles — sums off by ~1e-5 relative and different
    from run to run, while row counts stay exact. Repeats must agree bit for bit and match a torch
    fp64 restatement (sum order differs: 1e-12 relative)."""
    from velox_b200.kernels import FusedScanAgg, FusedScanCompact
    from velox_b200.queries import Q1, Q6, Q14, Q14_PROBE_SIG, Q14_SCAN_SIG

    nparts = 50_000
    li = tpch.gen_lineitem(rows, nparts, seed=5, device="cuda")
    part = tpch.gen_part(nparts, seed=6, device="cuda")
    price, disc, tax, qty, ship = li["l_extendedprice"], li["l_discount"], li["l_tax"], li["l_quantity"], li["l_shipdate"]

    def close(a, b):
        return abs(a - b) <= 1e-12 * abs(b)

    # post-exchange probe pipeline (two 8-byte columns: the fastest consumer loop of all)
    q14 = Q14()
    slot_flags, join_min = q14._build(part["p_partkey"], part["p_type"])
    rev = (price * (1.0 - disc)).contiguous()
    promo = torch.tensor([s.startswith("PROMO") for s in tpch.PTYPE_DICT], device="cuda")[part["p_type"].long()]
    promo_by_key = torch.zeros(nparts + 2, dtype=torch.bool, device="cuda")
    promo_by_key[part["p_partkey"]] = promo
    want = (float(rev.sum()), float(rev[promo_by_key[li["l_partkey"]]].sum()))
    probe = FusedScanAgg(Q14_PROBE_SIG)
    seen = set()
    for _ in range(5):
        probe.reset()
        probe.add_batch([li["l_partkey"], rev], rows, pf=[0.0], join={"slot_flags": slot_flags, "min": join_min})
        got = tuple(probe.sums.cpu().tolist())
        seen.add(got)
        assert int(probe.counts.item()) == rows and close(got[0], want[0]) and close(got[1], want[1]), (got, want)
    assert len(seen) == 1

    # Q6, Q1 and the single-GPU Q14 pipelines
    m6 = (ship >= tpch.Q6_SHIP_LO) & (ship <= tpch.Q6_SHIP_HI) & (disc >= 0.05) & (disc <= 0.07) & (qty < 24.0)
    want6 = float((price * disc)[m6].sum())
    q6 = Q6()
    seen = set()
    for _ in range(5):
        q6.launch(li, rows)
        seen.add(q6.result())
        assert close(q6.result(), want6)
    assert len(seen) == 1

    m1 = ship < tpch.Q1_SHIPDATE_LT
    q1 = Q1()
    seen = set()
    for _ in range(3):
        q1.launch(li, rows)
        res = q1.result()
        seen.add(tuple(sorted((k, v) for k, v in res.items())))
        for (rf, ls), v in res.items():
            g = m1 & (li["l_returnflag"] == tpch.RETURNFLAG_DICT.index(rf)) & (li["l_linestatus"] == tpch.LINESTATUS_DICT.index(ls))
            assert v[7] == int(g.sum())
            assert close(v[0], float(qty[g].sum())) and close(v[1], float(price[g].sum()))
            assert close(v[2], float((price * (1.0 - disc))[g].sum())) and close(v[3], float(((price * (1.0 - disc)) * (1.0 + tax))[g].sum()))
    assert len(seen) == 1

    m14 = (ship >= tpch.Q14_SHIP_LO) & (ship <= tpch.Q14_SHIP_HI)
    tot = float(rev[m14].sum())
    want14 = 100.0 * float(rev[m14 & promo_by_key[li["l_partkey"]]].sum()) / tot
    seen = set()
    for _ in range(5):
        q14.launch(li, part, rows)
        seen.add(q14.result())
        assert close(q14.result(), want14)
    assert len(seen) == 1

    # scan-compact with every row passing: (key, value) pairs must come out as the same multiset
    li2 = dict(li)
    li2["l_shipdate"] = torch.full_like(ship, tpch.Q14_SHIP_LO)
    scan = FusedScanCompact(Q14_SCAN_SIG, rows + 1024)
    for _ in range(3):
        scan.run([li2["l_shipdate"], li["l_partkey"], price, disc], rows, pf=[1.0], pi=[tpch.Q14_SHIP_LO, tpch.Q14_SHIP_HI])
        n, (lk, rv) = scan.result([torch.int64, torch.float64])
        assert n == rows
        a = torch.zeros(nparts + 1, dtype=torch.float64, device="cuda").index_add_(0, lk, rv)
        b = torch.zeros(nparts + 1, dtype=torch.float64, device="cuda").index_add_(0, li["l_partkey"], rev)
        assert float(((a - b).abs() / b.abs().clamp_min(1.0)).max()) < 1e-12
        assert bool((torch.sort(lk).values == torch.sort(li["l_partkey"]).values).all())


@pytest.mark.parametrize("n", [0, 1, 4095, 4096, 100_003])
def test_partition_segments_match_oracle_partitioning(n):
    """Sync-free fixed-capacity partitioning (vb2k_partition_segments): partition of every row equals
    HashPartitionFunction on the CPU oracle (exec/HashPartitionFunction.cpp:113-116), rows keep their
    input order inside a segment, tails hold the sentinel, the device-side row count bounds the input,
    and an undersized segment raises the overflow flag."""
    from velox_b200.kernels import SENTINEL_KEY, partition_segments
    rng = np.random.default_rng(n + 1)
    keys = rng.integers(-2**62, 2**62, n)
    pay8 = rng.standard_normal(n)
    pay4 = rng.integers(0, 1000, n).astype(np.int32)
    host_key = flat_vector(BIGINT, keys)
    tk = torch.from_numpy(keys).cuda() if n else torch.zeros(0, dtype=torch.int64, device="cuda")
    t8 = torch.from_numpy(pay8).cuda() if n else torch.zeros(0, dtype=torch.float64, device="cuda")
    t4 = torch.from_numpy(pay4).cuda() if n else torch.zeros(0, dtype=torch.int32, device="cuda")
    for parts in (1, 2, 8, 13):
        want = pyoracle.partition([host_key], parts) if n else np.zeros(0, dtype=np.uint32)
        counts = np.bincount(want, minlength=parts) if n else np.zeros(parts, dtype=np.int64)
        segcap = int(max(64, counts.max() + 3))
        flag = torch.zeros(2, dtype=torch.int32, device="cuda")
        # slack rows after the logical end must be ignored through the device-side row count
        pad = 37
        tk2 = torch.cat([tk, torch.full((pad,), 5, dtype=torch.int64, device="cuda")])
        t82 = torch.cat([t8, torch.zeros(pad, dtype=torch.float64, device="cuda")])
        t42 = torch.cat([t4, torch.zeros(pad, dtype=torch.int32, device="cuda")])
        rows_dev = torch.tensor([n], dtype=torch.int64, device="cuda")
        sk, (s8, s4), cnt = partition_segments(tk2, [t82, t42], n + pad, rows_dev, parts, segcap, flag)
        assert cnt.cpu().tolist() == counts.tolist() and int(flag[0].item()) == 0
        sk, s8, s4 = sk.cpu().numpy(), s8.cpu().numpy(), s4.cpu().numpy()
        for p in range(parts):
            rows = np.nonzero(want == p)[0] if n else np.zeros(0, dtype=np.int64)
            seg = slice(p * segcap, p * segcap + len(rows))
            assert np.array_equal(sk[seg], keys[rows]) and np.array_equal(s8[seg], pay8[rows]) and np.array_equal(s4[seg], pay4[rows])
            assert np.all(sk[p * segcap + len(rows):(p + 1) * segcap] == SENTINEL_KEY)
        if n > parts and counts.max() > 1:
            flag.zero_()
            partition_segments(tk, [t8], n, None, parts, int(counts.max()) - 1, flag)
            assert int(flag[0].item()) == 1


def test_q1_q6_q14_oracle_parity_at_24m_rows():
    """Oracle-vs-GPU at a size where every TMA stage of every resident block is recycled hundreds of
    times and the persistent grid wraps (24 M rows): Q1 / Q6 / Q14 through the operator API, fused and
    late-materialisation strategies, against the multi-threaded CPU oracle on the same rows."""
    import bench
    from oracle import pyoracle
    from velox_b200 import tpch
    from velox_b200.task import run_plan
    n, nparts = 24_000_000, 400_000
    li = tpch.gen_lineitem(n, nparts, seed=123, device="cpu")
    part = tpch.gen_part(nparts, seed=5, device="cpu")
    rv1, rv14, pt = bench.host_tables(li, part, n)
    p1, p14 = bench.plans(rv1, rv14, pt)
    threads = min(32, os.cpu_count() or 1)
    w1 = pyoracle.run_plan(p1, [rv1], threads=threads, batch_rows=100_000)
    w14 = pyoracle.run_plan(p14, [rv14, pt], threads=threads, batch_rows=100_000)
    for cfg in ({"b200.late_materialization": "false"}, {}):
        g1, s1 = run_plan(p1, [rv1], config=cfg)
        g14, s14 = run_plan(p14, [rv14, pt], config=cfg)
        par = bench.parity(g1, g14, w1, w14)
        assert par["exact_columns_ok"] and par["fp_max_rel_err"] <= 1e-11, par
        assert sum(v for k, v in s1.items() if k.endswith("b200.fusedBatches")) == 1
    assert sum(v for k, v in s14.items() if k.endswith("b200.selectiveBatches")) == 1
