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
