"""World-size-2 check of the hash-partitioned exchange logic on CPU (gloo): partition ids follow
HashPartitionFunction (hash % world, velox/exec/HashPartitionFunction.cpp:113-116); after one
all-to-all every rank owns exactly the keys that hash to it, both join sides meet on the same rank,
and the per-rank partial results add up to the single-process answer. The GPU path
(velox_b200/queries.py Q14._launch_partitioned) runs the same steps with the device kernels + NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import pyoracle
    from velox_b200 import tpch
    from velox_b200.vector import BIGINT, flat_vector
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, nparts = 40_000, 1000
    li = {k: v.numpy() for k, v in tpch.gen_lineitem(n, nparts, seed=42 + rank, device="cpu").items()}
    part = {k: v.numpy() for k, v in tpch.gen_part(nparts, seed=43, device="cpu").items()}
    p0, p1 = nparts * rank // world, nparts * (rank + 1) // world

    def all_to_all(recv, send):
        """gloo has no alltoall: pairwise blocking send/recv in a deadlock-free order."""
        for p in range(world):
            if p == rank:
                recv[p].copy_(send[p])
            elif rank < p:
                dist.send(send[p], p)
                dist.recv(recv[p], p)
            else:
                dist.recv(recv[p], p)
                dist.send(send[p], p)

    def exchange(keys, payload):
        ids = pyoracle.partition([flat_vector(BIGINT, keys)], world).astype(np.int64)
        order = np.argsort(ids, kind="stable")
        counts = np.bincount(ids, minlength=world)
        send_k = [torch.from_numpy(np.ascontiguousarray(keys[order][counts[:p].sum():counts[:p + 1].sum()])) for p in range(world)]
        send_p = [torch.from_numpy(np.ascontiguousarray(payload[order][counts[:p].sum():counts[:p + 1].sum()])) for p in range(world)]
        rc = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
        all_to_all(rc, [torch.tensor([int(c)]) for c in counts])
        recv_k = [torch.empty(int(c.item()), dtype=send_k[0].dtype) for c in rc]
        recv_p = [torch.empty(int(c.item()), dtype=send_p[0].dtype) for c in rc]
        all_to_all(recv_k, send_k)
        all_to_all(recv_p, send_p)
        return torch.cat(recv_k).numpy(), torch.cat(recv_p).numpy()

    pk, pt = exchange(part["p_partkey"][p0:p1], part["p_type"][p0:p1])
    m = (li["l_shipdate"] >= tpch.Q14_SHIP_LO) & (li["l_shipdate"] <= tpch.Q14_SHIP_HI)
    rev = li["l_extendedprice"][m] * (1.0 - li["l_discount"][m])
    lk, lrev = exchange(li["l_partkey"][m], rev)
    # every received key belongs to this rank
    assert (pyoracle.partition([flat_vector(BIGINT, pk)], world) == rank).all()
    assert len(lk) == 0 or (pyoracle.partition([flat_vector(BIGINT, lk)], world) == rank).all()
    promo_by_key = dict(zip(pk.tolist(), (np.array([s.startswith("PROMO") for s in tpch.PTYPE_DICT])[pt]).tolist()))
    total = float(lrev.sum())
    promo = float(sum(r for k, r in zip(lk.tolist(), lrev.tolist()) if promo_by_key[k]))  # KeyError = sides did not meet
    t = torch.tensor([total, promo, float(len(lk)), float(m.sum())], dtype=torch.float64)
    dist.all_reduce(t)
    if rank == 0:
        out.put(t.tolist())
    dist.destroy_process_group()


def test_partitioned_join_exchange_world2():
    from velox_b200 import tpch
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    # single-process answer over the union of both ranks' shards
    nparts = 1000
    part = {k: v.numpy() for k, v in tpch.gen_part(nparts, seed=43, device="cpu").items()}
    promo_flag = np.array([s.startswith("PROMO") for s in tpch.PTYPE_DICT])[part["p_type"]]
    total = promo = rows = 0.0
    for r in range(world):
        li = {k: v.numpy() for k, v in tpch.gen_lineitem(40_000, nparts, seed=42 + r, device="cpu").items()}
        m = (li["l_shipdate"] >= tpch.Q14_SHIP_LO) & (li["l_shipdate"] <= tpch.Q14_SHIP_HI)
        rev = li["l_extendedprice"][m] * (1.0 - li["l_discount"][m])
        total += rev.sum()
        promo += rev[promo_flag[li["l_partkey"][m] - 1]].sum()
        rows += m.sum()
    assert got[2] == rows == got[3]           # no row lost or duplicated by the exchange
    assert got[0] == pytest.approx(total, rel=1e-12) and got[1] == pytest.approx(promo, rel=1e-12)


# ---- planned (sync-free) exchange: fixed-capacity segments padded with a sentinel key ----------------
SENTINEL = -0x7F7F7F7F7F7F7F80  # VB2_SENTINEL_KEY


def _planned_worker(rank, world, port, out, undersized):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import pyoracle
    from velox_b200 import tpch
    from velox_b200.vector import BIGINT, flat_vector
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n, nparts = 40_000, 1000
    li = {k: v.numpy() for k, v in tpch.gen_lineitem(n, nparts, seed=42 + rank, device="cpu").items()}
    part = {k: v.numpy() for k, v in tpch.gen_part(nparts, seed=43, device="cpu").items()}
    p0, p1 = nparts * rank // world, nparts * (rank + 1) // world

    def all_to_all(recv, send):
        for p in range(world):
            if p == rank:
                recv[p].copy_(send[p])
            elif rank < p:
                dist.send(send[p], p)
                dist.recv(recv[p], p)
            else:
                dist.recv(recv[p], p)
                dist.send(send[p], p)

    def segments(keys, payload, segcap):
        """numpy restatement of vb2k_partition_segments: stable rank inside the destination segment,
        sentinel tails, overflow flag when a partition exceeds the planned capacity."""
        ids = pyoracle.partition([flat_vector(BIGINT, keys)], world).astype(np.int64) if len(keys) else np.zeros(0, dtype=np.int64)
        seg_k = np.full(world * segcap, SENTINEL, dtype=np.int64)
        seg_p = np.zeros(world * segcap, dtype=payload.dtype)
        overflow = False
        for p in range(world):
            rows = np.nonzero(ids == p)[0]
            if len(rows) > segcap:
                overflow = True
                rows = rows[:segcap]
            seg_k[p * segcap:p * segcap + len(rows)] = keys[rows]
            seg_p[p * segcap:p * segcap + len(rows)] = payload[rows]
        return seg_k, seg_p, overflow

    def planned_exchange(keys, payload, segcap):
        sk, sp, ov = segments(keys, payload, segcap)
        send_k = [torch.from_numpy(sk[p * segcap:(p + 1) * segcap].copy()) for p in range(world)]
        send_p = [torch.from_numpy(sp[p * segcap:(p + 1) * segcap].copy()) for p in range(world)]
        recv_k = [torch.empty(segcap, dtype=torch.int64) for _ in range(world)]
        recv_p = [torch.empty(segcap, dtype=send_p[0].dtype) for _ in range(world)]
        all_to_all(recv_k, send_k)   # equal sizes on every rank: no count exchange, nothing to wait for
        all_to_all(recv_p, send_p)
        return torch.cat(recv_k).numpy(), torch.cat(recv_p).numpy(), ov

    # statistics of a planning run, identical on every rank (the GPU path max-all-reduces them)
    m = (li["l_shipdate"] >= tpch.Q14_SHIP_LO) & (li["l_shipdate"] <= tpch.Q14_SHIP_HI)
    rev = li["l_extendedprice"][m] * (1.0 - li["l_discount"][m])
    lkeys = li["l_partkey"][m]
    stats = torch.tensor([np.bincount(pyoracle.partition([flat_vector(BIGINT, lkeys)], world), minlength=world).max() if len(lkeys) else 0,
                          np.bincount(pyoracle.partition([flat_vector(BIGINT, part["p_partkey"][p0:p1])], world), minlength=world).max()], dtype=torch.int64)
    dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    li_seg = max(8, int(stats[0]) * 5 // 4) if not undersized else max(1, int(stats[0]) // 2)
    part_seg = max(8, int(stats[1]) * 5 // 4)

    pk, pt, ov1 = planned_exchange(part["p_partkey"][p0:p1], part["p_type"][p0:p1], part_seg)
    lk, lrev, ov2 = planned_exchange(lkeys, rev, li_seg)
    flag = torch.tensor([int(ov1 or ov2)], dtype=torch.int64)
    dist.all_reduce(flag)  # a broken planning assumption anywhere invalidates the run everywhere
    live_p, live_l = pk != SENTINEL, lk != SENTINEL
    assert (pyoracle.partition([flat_vector(BIGINT, pk[live_p])], world) == rank).all()
    promo_by_key = dict(zip(pk[live_p].tolist(), (np.array([s.startswith("PROMO") for s in tpch.PTYPE_DICT])[pt[live_p]]).tolist()))
    total = float(lrev[live_l].sum())
    promo = float(sum(r for k, r in zip(lk[live_l].tolist(), lrev[live_l].tolist()) if promo_by_key.get(k, False)))
    t = torch.tensor([total, promo, float(live_l.sum()), float(m.sum()), float(flag.item())], dtype=torch.float64)
    dist.all_reduce(t[:4])
    if rank == 0:
        out.put(t.tolist())
    dist.destroy_process_group()


@pytest.mark.parametrize("undersized", [False, True])
def test_planned_exchange_world2(undersized):
    """Fixed-capacity sentinel-padded segments, no count exchange (Q14._launch_planned on the GPU):
    with planned sizes + head-room nothing is lost; with an undersized plan rows are dropped AND
    the overflow flag is raised on every rank, which is what makes the query rerun as a planning run."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_planned_worker, args=(r, world, port, q, undersized)) for r in range(world)]
    for p in procs:
        p.start()
    total, promo, joined, scanned, flag = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    if undersized:
        assert flag > 0 and joined < scanned
    else:
        assert flag == 0 and joined == scanned
