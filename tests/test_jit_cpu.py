"""Expression JIT code generation, checked without a GPU: the generated CUDA source of a program
that uses every opcode / type / encoding combination must compile for sm_100a with NVRTC (the
kernels themselves run in the GPU suite, where every operator test runs with the JIT and with the
interpreter)."""
import torch  # noqa: F401  (first: our library binds to the CUDA libraries torch loads)
import ctypes as C

from velox_b200._lib import lib
from velox_b200.kernels import CColumn, Const, Instr, Output, Program

B, I, BI, D, V = 0, 3, 4, 6, 7


def col(t, enc, nulls=False, dict_nulls=False):
    c = CColumn()
    c.type, c.encoding, c.size = t, enc, 100
    c.values, c.aux = 0x1000, 0x5000  # never dereferenced: nothing is launched
    c.nulls = 0x2000 if nulls else None
    c.indices = 0x3000 if enc == 1 else None
    c.dict_size = 10
    c.dict_nulls = 0x4000 if dict_nulls else None
    return c


def compiles(prog, cols, filt, outs):
    L = lib()
    L.vb2k_expression_jit_compiles.restype = C.c_int32
    buf = C.create_string_buffer(1 << 18)
    arr = (CColumn * len(cols))(*cols)
    oa = (Output * max(1, len(outs)))(*outs)
    rc = L.vb2k_expression_jit_compiles(C.byref(prog), arr, len(cols), 1 if filt else 0, oa, len(outs), buf, len(buf))
    return rc, buf.value.decode(errors="replace")


def test_every_opcode_compiles_for_sm_100a():
    cols = [col(BI, 0, True), col(D, 1, True, True), col(I, 2, True), col(B, 0), col(V, 1, False, True), col(D, 0)]
    ins = []

    def add(op, t, d, a=0, b=0, c=0):
        ins.append(Instr(op, t, d, a, b, c))

    add(1, BI, 0, 0); add(1, D, 1, 1); add(1, I, 2, 2); add(1, B, 3, 3); add(1, D, 4, 5)   # LOAD: flat / dictionary / constant
    add(2, BI, 5, 0); add(2, D, 6, 1); add(24, D, 7)                                       # CONST, NULL
    for op in (3, 4, 5, 6, 7):                                                             # + - * / % on BIGINT, INTEGER, DOUBLE
        add(op, BI, 8, 0, 5); add(op, I, 9, 2, 2); add(op, D, 10, 1, 6)
    add(8, BI, 8, 0); add(8, I, 9, 2); add(8, D, 10, 1)                                    # NEG
    for op in range(9, 15):                                                                # comparisons
        add(op, D, 11, 1, 6); add(op, BI, 12, 0, 5)
    add(15, D, 11, 1, 6, 4); add(15, BI, 12, 0, 5, 5)                                      # BETWEEN
    add(16, B, 13, 11, 12); add(17, B, 13, 11, 12); add(18, B, 14, 13); add(19, B, 14, 7)  # AND OR NOT IS_NULL
    add(20, D, 15, 13, 1, 6); add(20, D, 15, 13, 1, -1)                                    # CASE with / without ELSE
    for to, frm, src in ((D, BI, 0), (BI, D, 1), (I, BI, 0), (BI, I, 2), (I, D, 1), (BI, BI, 0)):
        add(21, to, 16, src, frm)                                                          # CAST
    add(22, B, 17, 4, 2); add(23, B, 17, 4, 2, 3)                                          # LIKE, string compare
    add(2, D, 18, 3)                                                                       # NULL constant
    consts = [Const(BI, 0, 5, 0.0, None, 0, 0), Const(D, 0, 0, 1.5, None, 0, 0), Const(V, 0, 0, 0.0, 0x6000, 3, 0), Const(D, 1, 0, 0.0, None, 0, 0)]
    ia, ca = (Instr * len(ins))(*ins), (Const * len(consts))(*consts)
    prog = Program(ia, len(ins), len(ins), 13, 19, ca, len(consts), 0)
    rc, text = compiles(prog, cols, True, [])
    assert rc == 1, text[:4000]
    outs = [Output(10, D, 0x7000, 0x8000), Output(9, I, 0x7000, 0x8000), Output(13, B, 0x7000, 0x8000), Output(16, BI, 0x7000, 0x8000)]
    rc, text = compiles(prog, cols, False, outs)
    assert rc == 1, text[:4000]
    # the generated text is specialised: no validity lookup for the column without NULLs
    assert "a.cols[5].nulls" not in text and "a.cols[0].nulls" in text


def test_unsupported_program_falls_back():
    cols = [col(BI, 0)]
    ia = (Instr * 1)(Instr(99, BI, 0, 0, 0, 0))  # unknown opcode
    ca = (Const * 1)(Const(BI, 0, 0, 0.0, None, 0, 0))
    prog = Program(ia, 1, 0, -1, 1, ca, 1, 0)
    rc, _ = compiles(prog, cols, False, [Output(0, BI, 0x7000, 0x8000)])
    assert rc == 0  # the interpreter (which ignores unknown opcodes the same way it always has) runs instead


def test_tpch_plans_jit_compile():
    """Every Filter / Project node of the TPC-H Q1, Q6 and Q14 plans (and a plan exercising CASE,
    CAST, LIKE, checked integer arithmetic and three-valued logic) goes through the expression
    compiler and yields kernels that NVRTC compiles for sm_100a — none is left to the interpreter."""
    import numpy as np
    from velox_b200 import tpch
    from velox_b200.plan import PlanBuilder
    from velox_b200.vector import BIGINT, BOOLEAN, DOUBLE, INTEGER, VARCHAR

    li_names = ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate", "l_partkey"]
    li_types = [VARCHAR, VARCHAR, DOUBLE, DOUBLE, DOUBLE, DOUBLE, INTEGER, BIGINT]
    q1 = (PlanBuilder().values(li_names, li_types).filter("l_shipdate < '1998-09-03'::DATE")
          .project(["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_extendedprice * (1.0 - l_discount) AS a",
                    "l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS b", "l_discount"])
          .partialAggregation(["l_returnflag", "l_linestatus"], ["sum(l_quantity)", "sum(a)", "sum(b)", "avg(l_discount)", "count(0)"])
          .localPartition([]).finalAggregation().planNode())
    q6 = (PlanBuilder().values(li_names, li_types)
          .filter("l_shipdate between '1994-01-01'::DATE and '1994-12-31'::DATE and l_discount between 0.05 and 0.07 and l_quantity < 24.0")
          .project(["l_extendedprice * l_discount"]).partialAggregation([], ["sum(p0)"]).localPartition([]).finalAggregation().planNode())
    build = PlanBuilder().values(["p_partkey", "p_type"], [BIGINT, VARCHAR], source=1)
    q14 = (PlanBuilder().values(li_names, li_types, source=0).filter("l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE")
           .project(["l_extendedprice * (1.0 - l_discount) as part_revenue", "l_shipdate", "l_partkey"])
           .hashJoin(["l_partkey"], ["p_partkey"], build, "", ["part_revenue", "p_type"])
           .project(["(CASE WHEN (p_type LIKE 'PROMO%') THEN part_revenue ELSE 0.0 END) as filter_revenue", "part_revenue"])
           .partialAggregation([], ["sum(part_revenue) as t", "sum(filter_revenue) as p"]).localPartition([]).finalAggregation()
           .project(["100.00 * p / t as promo_revenue"]).planNode())
    misc = (PlanBuilder().values(["i", "j", "d", "s", "b"], [BIGINT, INTEGER, DOUBLE, VARCHAR, BOOLEAN])
            .filter("(i + 1 > j * 2 OR b) AND NOT (d IS NULL) AND s LIKE '%x_'")
            .project(["CAST(i AS DOUBLE) / d AS q", "CASE WHEN j > 3 THEN i - j ELSE i % 7 END AS c", "-d AS n", "CAST(d AS BIGINT) AS t", "s < 'm' AS lt"]).planNode())
    L = lib()
    for name, plan, min_programs in (("q1", q1, 2), ("q6", q6, 2), ("q14", q14, 4), ("misc", misc, 2)):
        progs, jit, total = C.c_int32(0), C.c_int32(0), C.c_int32(0)
        err = C.create_string_buffer(2048)
        rc = L.vb2_plan_jit_report(plan.sexpr.encode(), C.byref(progs), C.byref(jit), C.byref(total), err, 2048)
        assert rc == 0, (name, err.value.decode())
        assert progs.value >= min_programs and total.value >= progs.value, (name, progs.value, total.value)
        assert jit.value == total.value, f"{name}: {total.value - jit.value} of {total.value} expression kernels fall back to the interpreter"


def test_pipeline_jit_compiles_template_instantiations():
    """fused_jit.cu (no GPU needed): signatures the planner prints are parsed back into the expression
    template they describe and every kernel family compiles for sm_100a with NVRTC."""
    import ctypes as C
    from velox_b200 import tpch
    from velox_b200._lib import lib
    L = lib()
    L.vb2k_pipeline_jit_compiles.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
    q6x = tpch.Q6_SIG.replace("lt(f2,pf2))", "lt(f2,pf2),gt(f3,pf3))")
    cases = [(q6x, 0, 1, 0), (q6x, 2, 0, 0), (q6x, 3, 1, 0), (tpch.Q1_SIG, 0, 0, 0), (tpch.Q14_SIG, 3, 1, 0), (tpch.Q14_SIG, 1, 1, 0),
             ("F:between(i0,pi0,pi1);P:", 2, 0, 0), ("F:lt(f0,pf0);P:divide(f1,plus(pf1,f0))", 0, 4, 1)]
    for sig, kind, groups, key64 in cases:
        buf = C.create_string_buffer(4000)
        assert L.vb2k_pipeline_jit_compiles(sig.encode(), kind, groups, key64, buf, 4000) == 1, (sig, kind, buf.value.decode()[:1500])
        assert "<" in buf.value.decode()
    buf = C.create_string_buffer(400)
    assert L.vb2k_pipeline_jit_compiles(b"F:or(lt(f0,pf0),lt(f1,pf1));P:f0", 0, 1, 0, buf, 400) == 0  # outside the template grammar
