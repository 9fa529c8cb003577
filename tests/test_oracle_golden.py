"""Pins the CPU oracle before it is trusted as the checker (CPU only, no GPU):
  * scalar kernels against known-answer vectors transcribed from the reference's own tests
    (tests/golden/scalar_vectors.json — each entry cites its file:line);
  * hash mixers against the published algorithms (independent pure-Python restatements, CRC-32C
    anchored on its standard check value) and the symbolic relations of VectorHasherTest;
  * whole-plan results (Q1 / Q6 / Q14 shapes, joins, group by) against numpy / pyarrow
    re-computations — the reference pins these through DuckDB, which is absent here."""
import json
import math
import os
import struct

import numpy as np
import pytest

from oracle import pyoracle
from velox_b200 import tpch
from velox_b200.plan import PlanBuilder
from velox_b200.vector import (BIGINT, BOOLEAN, DOUBLE, INTEGER, VARCHAR, constant_vector, dictionary_vector, flat_vector,
                               row_vector)

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "scalar_vectors.json")))
M64 = (1 << 64) - 1


def f(x):
    return {"nan": float("nan"), "snan": struct.unpack("<d", struct.pack("<Q", 0x7ff0000000000001))[0], "inf": float("inf")}.get(x, x) if isinstance(x, str) else x


def project(rv, exprs):
    return pyoracle.run_plan(PlanBuilder().values(rv.names, rv.types).project(exprs).planNode(), [rv]).rows()


def test_divide_known_answers():
    g = GOLD["divide_bigint"]
    rv = row_vector(["a", "b"], [flat_vector(BIGINT, g["a"]), flat_vector(BIGINT, g["b"])])
    assert [r[0] for r in project(rv, ["a / b"])] == g["expected"]
    g = GOLD["divide_integer"]
    rv = row_vector(["a", "b"], [flat_vector(INTEGER, g["a"]), flat_vector(INTEGER, g["b"])])
    assert [r[0] for r in project(rv, ["a / b"])] == g["expected"]
    g = GOLD["divide_double"]
    rv = row_vector(["a", "b"], [flat_vector(DOUBLE, g["a"]), flat_vector(DOUBLE, g["b"])])
    got = [r[0] for r in project(rv, ["a / b"])]
    for x, y in zip(got, [f(v) for v in g["expected"]]):
        assert (math.isnan(x) and math.isnan(y)) or x == y


def test_arithmetic_errors():
    for a, b in GOLD["divide_by_zero_errors"]["cases"]:
        rv = row_vector(["a", "b"], [flat_vector(INTEGER, [a]), flat_vector(INTEGER, [b])])
        with pytest.raises(pyoracle.OracleUserError):
            project(rv, ["a / b"])
    rv = row_vector(["a", "b"], [flat_vector(INTEGER, [-2**31]), flat_vector(INTEGER, [-1])])
    for e in ("a / b", "a * b"):
        with pytest.raises(pyoracle.OracleUserError):
            project(rv, [e])
    L = pyoracle.lib()
    import ctypes as C
    out = C.c_int64()
    assert L.orc_checked_i64(0, 2**63 - 1, 1, C.byref(out)) == 1
    assert L.orc_checked_i64(1, -2**63, 1, C.byref(out)) == 1
    assert L.orc_checked_i64(2, 2**32, 2**31, C.byref(out)) == 1
    assert L.orc_checked_i64(0, 5, 7, C.byref(out)) == 0 and out.value == 12


def test_nan_ordering():
    L = pyoracle.lib()
    for c in GOLD["nan_ordering"]["cases"]:
        a, b = f(c["a"]), f(c["b"])
        got = [bool(L.orc_compare_f64(op, a, b)) for op in (2, 3, 0, 1)]  # gt gte lt lte
        assert got == c["expected"], c
    nan = float("nan")
    assert L.orc_compare_f64(4, nan, nan) == 1 and L.orc_compare_f64(5, nan, nan) == 0  # NaN = NaN


# ---- hashing -----------------------------------------------------------------------------------
def py_twang(key):
    key = (~key + (key << 21)) & M64
    key ^= key >> 24
    key = (key + (key << 3) + (key << 8)) & M64
    key ^= key >> 14
    key = (key + (key << 2) + (key << 4)) & M64
    key ^= key >> 28
    key = (key + (key << 31)) & M64
    return key


def py_jenkins(key):
    m = (1 << 32) - 1
    key = (key + (key << 12)) & m
    key ^= key >> 22
    key = (key + (key << 4)) & m
    key ^= key >> 9
    key = (key + (key << 10)) & m
    key ^= key >> 2
    key = (key + (key << 7)) & m
    key = (key + (key << 12)) & m
    return key


def py_hash_mix(upper, lower):
    k = 0x9ddfea08eb382d69
    a = ((lower ^ upper) * k) & M64
    a ^= a >> 47
    b = ((upper ^ a) * k) & M64
    b ^= b >> 47
    return (b * k) & M64


def py_crc32c_update(crc, data: bytes):
    for byte in data:
        crc ^= byte
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 if crc & 1 else 0)
    return crc


def test_crc32c_and_hash_bytes():
    assert (py_crc32c_update(0xFFFFFFFF, b"123456789") ^ 0xFFFFFFFF) == int(GOLD["crc32c_check"]["expected"], 16)
    L = pyoracle.lib()
    for s in [b"", b"a", b"abc", b"PROMO", b"1234567"]:  # the < 8 byte branch of bits::hashBytes
        word = int.from_bytes(s.ljust(8, b"\0"), "little")
        crc = py_crc32c_update(1, word.to_bytes(8, "little"))
        crc2 = py_crc32c_update(1, (word >> 32).to_bytes(8, "little"))
        assert L.orc_hash_bytes(1, s, len(s)) == (crc | (crc2 << 32))
    k = 0x9ddfea08eb382d69
    for s in [b"12345678", b"PROMO BURNISHED COPPER", b"x" * 24, b"y" * 25, b"z" * 40, b"w" * 16, b"v" * 17]:
        a0, a1, a2 = 1, (1 << 32) & M64, 1 >> 16
        p, togo = 0, len(s)
        w = lambda i: s[p + 8 * i:p + 8 * i + 8]
        while togo >= 24:
            a0 = py_crc32c_update(a0 & 0xFFFFFFFF, w(0)); a1 = py_crc32c_update(a1 & 0xFFFFFFFF, w(1)); a2 = py_crc32c_update(a2 & 0xFFFFFFFF, w(2))
            p += 24; togo -= 24
        if togo > 16:
            a0 = py_crc32c_update(a0 & 0xFFFFFFFF, w(0)); a1 = py_crc32c_update(a1 & 0xFFFFFFFF, w(1))
            a2 = py_crc32c_update(a2 & 0xFFFFFFFF, s[p + 16:p + togo].ljust(8, b"\0"))
        elif togo > 8:
            a0 = py_crc32c_update(a0 & 0xFFFFFFFF, w(0))
            a1 = py_crc32c_update(a1 & 0xFFFFFFFF, s[p + 8:p + togo].ljust(8, b"\0"))
        elif togo > 0:
            a0 = py_crc32c_update(a0 & 0xFFFFFFFF, s[p:p + togo].ljust(8, b"\0"))
        assert L.orc_hash_bytes(1, s, len(s)) == (a0 ^ ((a1 * k) & M64) ^ ((a2 * k) & M64)), s


def test_mixers_match_published_algorithms():
    L = pyoracle.lib()
    rng = np.random.default_rng(0)
    for v in [0, 1, 7, 55, M64, 1 << 63] + rng.integers(0, 2**63, 200).tolist():
        assert L.orc_twang_mix64(v) == py_twang(v)
        assert L.orc_jenkins_rev_mix32(v & 0xFFFFFFFF) == py_jenkins(v & 0xFFFFFFFF)
        assert L.orc_hash_mix(v, (v * 31 + 7) & M64) == py_hash_mix(v, (v * 31 + 7) & M64)


def test_vector_hasher_relations():
    """VectorHasherTest.cpp:166-262: flat BIGINT == twang, NULL == kNullHash, mixing == hashMix,
    NaNs alike, +0 == -0, dictionary/constant encodings hash like their values."""
    vals = list(range(100))
    h = pyoracle.hash_columns([flat_vector(BIGINT, vals)])
    assert h.tolist() == [py_twang(v) for v in vals]
    h = pyoracle.hash_columns([flat_vector(BIGINT, [None, 7, None])])
    assert h.tolist() == [1, py_twang(7), 1]
    h2 = pyoracle.hash_columns([flat_vector(BIGINT, [7] * 3), flat_vector(BIGINT, [55] * 3)])
    assert h2.tolist() == [py_hash_mix(py_twang(7), py_twang(55))] * 3
    d = [f(x) for x in GOLD["vector_hasher"]["nans_input"]]
    h = pyoracle.hash_columns([flat_vector(DOUBLE, np.array(d))]).tolist()
    assert h[2] == h[3] and h[4] == h[5] == 0 and h[0] != h[1]
    assert h[0] == py_twang(struct.unpack("<Q", struct.pack("<d", 1.0))[0])
    i32 = pyoracle.hash_columns([flat_vector(INTEGER, [5, -5])]).tolist()
    assert i32 == [py_jenkins(5), py_jenkins((-5) & 0xFFFFFFFF)]
    # encodings
    base = ["apple", "banana", None]
    idx = [0, 1, 2, 1, 0]
    hd = pyoracle.hash_columns([dictionary_vector(VARCHAR, idx, base)])
    hf = pyoracle.hash_columns([flat_vector(VARCHAR, [base[i] for i in idx])])
    assert hd.tolist() == hf.tolist() and hd[2] == 1
    hc = pyoracle.hash_columns([constant_vector(BIGINT, 42, 4)])
    assert hc.tolist() == [py_twang(42)] * 4
    # partition = hash % P (HashPartitionFunction.cpp:113-116)
    p = pyoracle.partition([flat_vector(BIGINT, vals)], 8)
    assert p.tolist() == [py_twang(v) % 8 for v in vals]


# ---- whole plans against independent engines ----------------------------------------------------------
def _lineitem(n, seed, nparts=500):
    t = tpch.gen_lineitem(n, nparts, seed=seed, device="cpu")
    return {k: v.numpy() for k, v in t.items()}


def test_q6_q1_q14_against_numpy_and_pyarrow():
    import pyarrow as pa
    import pyarrow.compute as pc
    n, nparts = 60_000, 500
    h = _lineitem(n, 17, nparts)
    # Q6
    rv = row_vector(["l_shipdate", "l_extendedprice", "l_quantity", "l_discount"],
                    [flat_vector(INTEGER, h["l_shipdate"])] + [flat_vector(DOUBLE, h[c]) for c in ("l_extendedprice", "l_quantity", "l_discount")])
    plan = (PlanBuilder().values(rv.names, rv.types)
            .filter("l_shipdate between '1994-01-01'::DATE and '1994-12-31'::DATE and l_discount between 0.05 and 0.07 and l_quantity < 24.0")
            .project(["l_extendedprice * l_discount"]).singleAggregation([], ["sum(p0)"]).planNode())
    got = pyoracle.run_plan(plan, [rv]).rows()[0][0]
    m = ((h["l_shipdate"] >= tpch.Q6_SHIP_LO) & (h["l_shipdate"] <= tpch.Q6_SHIP_HI) & (h["l_discount"] >= 0.05) & (h["l_discount"] <= 0.07) & (h["l_quantity"] < 24))
    assert got == pytest.approx(float(np.sum(h["l_extendedprice"][m] * h["l_discount"][m])), rel=1e-12)
    # Q1 against pyarrow's group_by (an independent columnar engine)
    names = ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate"]
    cols = [dictionary_vector(VARCHAR, h["l_returnflag"], tpch.RETURNFLAG_DICT), dictionary_vector(VARCHAR, h["l_linestatus"], tpch.LINESTATUS_DICT)] + \
           [flat_vector(DOUBLE, h[c]) for c in names[2:6]] + [flat_vector(INTEGER, h["l_shipdate"])]
    rv = row_vector(names, cols)
    plan = (PlanBuilder().values(rv.names, rv.types).filter("l_shipdate < '1998-09-03'::DATE")
            .project(["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice * (1.0 - l_discount) AS dp",
                      "l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS ch", "l_discount"])
            .partialAggregation(["l_returnflag", "l_linestatus"], ["sum(l_quantity)", "sum(dp)", "sum(ch)", "avg(l_discount)", "count(0)"])
            .localPartition([]).finalAggregation().planNode())
    got = {(r[0], r[1]): r[2:] for r in pyoracle.run_plan(plan, [rv], threads=3).rows()}
    keep = h["l_shipdate"] < tpch.Q1_SHIPDATE_LT
    tbl = pa.table({"rf": np.array(tpch.RETURNFLAG_DICT)[h["l_returnflag"]][keep], "ls": np.array(tpch.LINESTATUS_DICT)[h["l_linestatus"]][keep],
                    "q": h["l_quantity"][keep], "dp": (h["l_extendedprice"] * (1 - h["l_discount"]))[keep],
                    "ch": (h["l_extendedprice"] * (1 - h["l_discount"]) * (1 + h["l_tax"]))[keep], "d": h["l_discount"][keep]})
    want = tbl.group_by(["rf", "ls"]).aggregate([("q", "sum"), ("dp", "sum"), ("ch", "sum"), ("d", "mean"), ("q", "count")]).to_pylist()
    assert len(want) == len(got)
    for w in want:
        g = got[(w["rf"], w["ls"])]
        assert g[4] == w["q_count"]
        for a, b in zip(g[:4], (w["q_sum"], w["dp_sum"], w["ch_sum"], w["d_mean"])):
            assert a == pytest.approx(b, rel=1e-11)
    # Q14 against numpy
    part = {k: v.numpy() for k, v in tpch.gen_part(nparts, seed=5).items()}
    li = row_vector(["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"],
                    [flat_vector(BIGINT, h["l_partkey"]), flat_vector(DOUBLE, h["l_extendedprice"]), flat_vector(DOUBLE, h["l_discount"]), flat_vector(INTEGER, h["l_shipdate"])])
    pt = row_vector(["p_partkey", "p_type"], [flat_vector(BIGINT, part["p_partkey"]), dictionary_vector(VARCHAR, part["p_type"], tpch.PTYPE_DICT)])
    b = PlanBuilder().values(pt.names, pt.types, source=1)
    plan = (PlanBuilder().values(li.names, li.types, source=0).filter("l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE")
            .project(["l_extendedprice * (1.0 - l_discount) as part_revenue", "l_shipdate", "l_partkey"])
            .hashJoin(["l_partkey"], ["p_partkey"], b, "", ["part_revenue", "p_type"])
            .project(["(CASE WHEN (p_type LIKE 'PROMO%') THEN part_revenue ELSE 0.0 END) as fr", "part_revenue"])
            .singleAggregation([], ["sum(part_revenue) as t", "sum(fr) as p"]).project(["100.00 * p / t"]).planNode())
    got = pyoracle.run_plan(plan, [li, pt], threads=2).rows()[0][0]
    m = (h["l_shipdate"] >= tpch.Q14_SHIP_LO) & (h["l_shipdate"] <= tpch.Q14_SHIP_HI)
    rev = h["l_extendedprice"][m] * (1 - h["l_discount"][m])
    promo = np.array([s.startswith("PROMO") for s in tpch.PTYPE_DICT])[part["p_type"][h["l_partkey"][m] - 1]]
    assert got == pytest.approx(100 * rev[promo].sum() / rev.sum(), rel=1e-12)


def test_join_and_groupby_against_pyarrow():
    import pyarrow as pa
    rng = np.random.default_rng(2)
    n, m = 3000, 400
    pk = [None if rng.random() < 0.05 else int(v) for v in rng.integers(0, 300, n)]
    bk = [None if rng.random() < 0.05 else int(v) for v in rng.integers(0, 350, m)]
    pv, bv = rng.integers(0, 1000, n), rng.integers(0, 1000, m)
    probe = row_vector(["pk", "pv"], [flat_vector(BIGINT, pk), flat_vector(BIGINT, pv)])
    build = row_vector(["bk", "bv"], [flat_vector(BIGINT, bk), flat_vector(BIGINT, bv)])
    bb = PlanBuilder().values(build.names, build.types, source=1)
    plan = PlanBuilder().values(probe.names, probe.types, source=0).hashJoin(["pk"], ["bk"], bb, "", ["pk", "pv", "bv"]).planNode()
    got = sorted(pyoracle.run_plan(plan, [probe, build]).rows())
    want = pa.table({"pk": pk, "pv": pv}).join(pa.table({"bk": bk, "bv": bv}), keys="pk", right_keys="bk", join_type="inner")
    want = sorted((r["pk"], r["pv"], r["bv"]) for r in want.to_pylist())
    assert got == want
    for jt, pa_jt in (("left", "left outer"), ("semi", "left semi"), ("anti", "left anti")):
        outs = ["pk", "pv", "bv"] if jt == "left" else ["pk", "pv"]
        plan = PlanBuilder().values(probe.names, probe.types, source=0).hashJoin(["pk"], ["bk"], PlanBuilder().values(build.names, build.types, source=1), "", outs, joinType=jt).planNode()
        got = pyoracle.run_plan(plan, [probe, build]).rows()
        w = pa.table({"pk": pk, "pv": pv}).join(pa.table({"bk": bk, "bv": bv}), keys="pk", right_keys="bk", join_type=pa_jt).to_pylist()
        key = lambda r: tuple((x is None, x) for x in r)
        assert sorted(got, key=key) == sorted((tuple(r[c] for c in outs) for r in w), key=key), jt
    # group by with null keys: null is a group (exec/GroupingSet.cpp:448-455)
    plan = PlanBuilder().values(probe.names, probe.types).singleAggregation(["pk"], ["sum(pv)", "count(0)"]).planNode()
    got = {r[0]: r[1:] for r in pyoracle.run_plan(plan, [probe], threads=4).rows()}
    want = pa.table({"pk": pk, "pv": pv}).group_by("pk").aggregate([("pv", "sum"), ("pv", "count")]).to_pylist()
    assert got == {r["pk"]: (r["pv_sum"], r["pv_count"]) for r in want}


def test_three_valued_logic_truth_table():
    t, fa, n = True, False, None
    a = [t, t, t, fa, fa, fa, n, n, n]
    b = [t, fa, n, t, fa, n, t, fa, n]
    rv = row_vector(["a", "b"], [flat_vector(BOOLEAN, a), flat_vector(BOOLEAN, b)])
    rows = project(rv, ["a and b", "a or b"])
    assert [r[0] for r in rows] == [t, fa, n, fa, fa, fa, n, fa, n]
    assert [r[1] for r in rows] == [t, t, t, t, fa, n, t, n, n]
