"""The expression JIT's generated kernels EXECUTED on the CPU. vb2k_expression_jit_compiles hands back the CUDA source it
generates for a program and column layout (expr_jit.cu; the text NVRTC compiles for sm_100a). That text is straight-line
per-row code plus warp ballots for the validity / selection bitmaps, so a small shim (thread indices as variables, a
two-pass warp ballot, atomicCAS, the __d*_rn intrinsics as plain IEEE operations without contraction) lets g++ compile
it and run it warp by warp. Its outputs are compared with the CPU oracle evaluating the same expressions as SQL — over
flat / dictionary / constant inputs with NULLs, NaN, checked-arithmetic errors, CASE, CAST, LIKE and three-valued logic.
No GPU needed; the GPU suite runs the same kernels on the device with both engines."""
import ctypes as C
import math
import os
import subprocess

import numpy as np
import pytest
import torch  # noqa: F401  (first: the library binds to the CUDA libraries torch loads)

from oracle import pyoracle
from velox_b200._lib import lib
from velox_b200.kernels import CColumn, Const, Instr, Output, Program
from velox_b200.plan import PlanBuilder
from velox_b200.vector import BIGINT, BOOLEAN, DOUBLE, INTEGER, VARCHAR, constant_vector, dictionary_vector, flat_vector, row_vector

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
T_BOOL, T_INT, T_BIG, T_DBL, T_STR = 0, 3, 4, 6, 7
NAN = float("nan")

SHIM = r"""
#include <cmath>
#include <cstring>
#include <vector>
namespace shim {
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static Dim3 threadIdx, blockIdx, gridDim;
static int phase = 0;            // 0: lanes record their ballot predicates, 1: ballots return the recorded masks
static size_t ballot_at = 0;     // index of the next ballot of the running lane (control flow around ballots is warp-uniform)
static std::vector<unsigned> masks;
static inline unsigned __ballot_sync(unsigned, bool p) {
  const size_t i = ballot_at++;
  if (phase == 0) {
    if (masks.size() <= i) masks.resize(i + 1, 0u);
    if (p) masks[i] |= 1u << (threadIdx.x & 31u);
    return 0u;
  }
  return masks[i];
}
static inline int atomicCAS(int* p, int expect, int value) { const int old = *p; if (old == expect) *p = value; return old; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline long long __mul64hi(long long a, long long b) { return static_cast<long long>((static_cast<__int128>(a) * b) >> 64); }
using std::isnan;
using std::round;
#define __global__
#define __device__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(x)
#define VB2_SIZEOF_COLUMN sizeof(vb2_column)
#define VB2_SIZEOF_CONST sizeof(vb2_const)
#define VB2_SIZEOF_OUTPUT sizeof(vb2_output)
#define VB2_SIZEOF_ARGS sizeof(JitArgs)
// ---- generated source ----
%s
// ---- driver: one block of 256 threads, warp by warp, two passes per warp ----
}  // namespace shim
extern "C" void run_on_host(const void* args) {
  using namespace shim;
  gridDim.x = 1;
  blockIdx.x = 0;
  const JitArgs& a = *static_cast<const JitArgs*>(args);
  for (unsigned warp = 0; warp < 8; ++warp) {
    masks.clear();
    for (phase = 0; phase < 2; ++phase)
      for (unsigned lane = 0; lane < 32; ++lane) {
        threadIdx.x = warp * 32 + lane;
        ballot_at = 0;
        vb2_jit(a);
      }
  }
}
"""


class JitArgs(C.Structure):
    _fields_ = [("cols", CColumn * 32), ("consts", Const * 32), ("outs", Output * 32), ("n", C.c_longlong), ("sel", C.c_void_p),
                ("sel_bits", C.c_void_p), ("error_flag", C.c_void_p)]


# ---- a tiny expression tree -> (vb2 program, SQL text) ------------------------------------------------
class Asm:
    """Registers are allocated one per node; columns: list of (name, vb2 type). LIKE / string compares take the column itself."""

    def __init__(self, columns):
        self.columns = columns
        self.ins, self.consts, self.keep = [], [], []
        self.reg = 0

    def _new(self):
        self.reg += 1
        return self.reg - 1

    def _const(self, typ, value):
        c = Const(typ, 0, 0, 0.0, None, 0, 0)
        if value is None:
            c.is_null = 1
        elif typ == T_DBL:
            c.d = float(value)
        elif typ == T_STR:
            buf = C.create_string_buffer(value.encode(), len(value.encode()))
            self.keep.append(buf)
            c.str, c.len = C.cast(buf, C.c_void_p).value, len(value.encode())
        else:
            c.i = int(value)
        self.consts.append(c)
        return len(self.consts) - 1

    def emit(self, e):
        """e: nested tuples. Returns (register, type, sql)."""
        op = e[0]
        if op == "col":
            idx = [n for n, _ in self.columns].index(e[1])
            typ = self.columns[idx][1]
            r = self._new()
            self.ins.append(Instr(1, typ, r, idx, 0, 0))
            return r, typ, e[1]
        if op == "lit":
            typ, v = e[1], e[2]
            r = self._new()
            self.ins.append(Instr(2, typ, r, self._const(typ, v), 0, 0))
            sql = {T_DBL: lambda x: repr(float(x)), T_BIG: lambda x: str(int(x)), T_INT: lambda x: f"cast({int(x)} as integer)",
                   T_BOOL: lambda x: "true" if x else "false"}[typ](v)
            return r, typ, sql
        if op in ("+", "-", "*", "/", "%"):
            (ra, ta, sa), (rb, _, sb) = self.emit(e[1]), self.emit(e[2])
            r = self._new()
            self.ins.append(Instr({"+": 3, "-": 4, "*": 5, "/": 6, "%": 7}[op], ta, r, ra, rb, 0))
            return r, ta, f"({sa} {op} {sb})"
        if op == "neg":
            ra, ta, sa = self.emit(e[1])
            r = self._new()
            self.ins.append(Instr(8, ta, r, ra, 0, 0))
            return r, ta, f"(-{sa})"
        if op in ("<", "<=", ">", ">=", "=", "<>"):
            (ra, ta, sa), (rb, _, sb) = self.emit(e[1]), self.emit(e[2])
            r = self._new()
            self.ins.append(Instr({"<": 9, "<=": 10, ">": 11, ">=": 12, "=": 13, "<>": 14}[op], ta, r, ra, rb, 0))
            return r, T_BOOL, f"({sa} {op} {sb})"
        if op == "between":
            (ra, ta, sa), (rb, _, sb), (rc, _, sc) = self.emit(e[1]), self.emit(e[2]), self.emit(e[3])
            r = self._new()
            self.ins.append(Instr(15, ta, r, ra, rb, rc))
            return r, T_BOOL, f"({sa} between {sb} and {sc})"
        if op in ("and", "or"):
            (ra, _, sa), (rb, _, sb) = self.emit(e[1]), self.emit(e[2])
            r = self._new()
            self.ins.append(Instr(16 if op == "and" else 17, T_BOOL, r, ra, rb, 0))
            return r, T_BOOL, f"({sa} {op} {sb})"
        if op == "not":
            ra, _, sa = self.emit(e[1])
            r = self._new()
            self.ins.append(Instr(18, T_BOOL, r, ra, 0, 0))
            return r, T_BOOL, f"(not {sa})"
        if op == "is_null":
            ra, _, sa = self.emit(e[1])
            r = self._new()
            self.ins.append(Instr(19, T_BOOL, r, ra, 0, 0))
            return r, T_BOOL, f"({sa} is null)"
        if op == "case":
            (rc, _, sc), (rt, tt, st) = self.emit(e[1]), self.emit(e[2])
            re_, se = -1, None
            if len(e) > 3:
                re_, _, se = self.emit(e[3])
            r = self._new()
            self.ins.append(Instr(20, tt, r, rc, rt, re_))
            return r, tt, f"(case when {sc} then {st}" + (f" else {se}" if se is not None else "") + " end)"
        if op == "cast":
            to = e[1]
            ra, ta, sa = self.emit(e[2])
            r = self._new()
            self.ins.append(Instr(21, to, r, ra, ta, 0))
            return r, to, f"cast({sa} as {({T_DBL: 'double', T_BIG: 'bigint', T_INT: 'integer', T_BOOL: 'boolean'})[to]})"
        if op == "like":
            idx = [n for n, _ in self.columns].index(e[1])
            r = self._new()
            self.ins.append(Instr(22, T_BOOL, r, idx, self._const(T_STR, e[2]), 0))
            return r, T_BOOL, f"({e[1]} like '{e[2]}')"
        if op == "strcmp":
            idx = [n for n, _ in self.columns].index(e[1])
            code = {"<": 0, "<=": 1, ">": 2, ">=": 3, "=": 4, "<>": 5}[e[2]]
            r = self._new()
            self.ins.append(Instr(23, T_BOOL, r, idx, self._const(T_STR, e[3]), code))
            return r, T_BOOL, f"({e[1]} {e[2]} '{e[3]}')"
        raise ValueError(op)


_cache = {}


def _host_kernel(source, tmp):
    if source in _cache:
        return _cache[source]
    i = len(_cache)
    src = tmp / f"jit{i}.cpp"
    src.write_text(SHIM % source)
    so = tmp / f"jit{i}.so"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-ffp-contract=off", "-w", "-I", os.path.join(ROOT, "velox_b200", "csrc"),
                           "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.run_on_host.argtypes = [C.c_void_p]
    _cache[source] = L
    return L


def run_projections(rv, exprs, tmp):
    """Evaluates `exprs` over the RowVector with the JIT-generated kernel on the host. Returns (columns as python lists with
    None for NULL, error_flag, sqls)."""
    asm = Asm([(n, c.type) for n, c in zip(rv.names, rv.columns)])
    regs = [asm.emit(e) for e in exprs]
    cols = [c.to_c() for c in rv.columns]  # host buffers in the C layout (validity bitmaps, bit-packed BOOLEAN); the Column objects keep them alive
    n = rv.size
    ia, ca = (Instr * len(asm.ins))(*asm.ins), (Const * max(1, len(asm.consts)))(*asm.consts)
    prog = Program(ia, len(asm.ins), 0, -1, asm.reg, ca, len(asm.consts), 0)
    outs, bufs = [], []
    for r, t, _ in regs:
        width = {T_DBL: 8, T_BIG: 8, T_INT: 4, T_BOOL: 1}[t]
        vals = np.zeros(n * width + 8, dtype=np.uint8)
        nulls = np.zeros((n + 31) // 32 + 2, dtype=np.uint32)
        bufs.append((vals, nulls, t))
        outs.append(Output(r, t, vals.ctypes.data, nulls.ctypes.data))
    L = lib()
    L.vb2k_expression_jit_compiles.restype = C.c_int32
    buf = C.create_string_buffer(1 << 18)
    arr = (CColumn * len(cols))(*cols)
    oa = (Output * len(outs))(*outs)
    rc = L.vb2k_expression_jit_compiles(C.byref(prog), arr, len(cols), 0, oa, len(outs), buf, len(buf))
    assert rc == 1, buf.value.decode(errors="replace")[:3000]
    K = _host_kernel(buf.value.decode(), tmp)
    a = JitArgs()
    for i, c in enumerate(cols):
        a.cols[i] = c
    for i, c in enumerate(asm.consts):
        a.consts[i] = c
    for i, o in enumerate(outs):
        a.outs[i] = o
    a.n = n
    err = np.zeros(2, dtype=np.int32)
    a.error_flag = err.ctypes.data
    K.run_on_host(C.byref(a))
    result = []
    for vals, nulls, t in bufs:
        dt = {T_DBL: np.float64, T_BIG: np.int64, T_INT: np.int32, T_BOOL: np.uint8}[t]
        v = vals[: n * np.dtype(dt).itemsize].view(dt)
        valid = [(int(nulls[i >> 5]) >> (i & 31)) & 1 for i in range(n)]
        result.append([None if not ok else (bool(x) if t == T_BOOL else (float(x) if t == T_DBL else int(x))) for x, ok in zip(v.tolist(), valid)])
    return result, int(err[0]), [s for _, _, s in regs]


def same(a, b):
    if a is None or b is None:
        return a is None and b is None
    if isinstance(b, float) or isinstance(a, float):
        return (math.isnan(a) and math.isnan(b)) or a == b
    return a == b


@pytest.fixture(scope="module")
def tmp(tmp_path_factory):
    return tmp_path_factory.mktemp("jit_on_host")


def table(n=300, seed=0):
    rng = np.random.default_rng(seed)

    def maybe(v, p=0.12):
        return [None if rng.random() < p else x for x in v]

    return row_vector(
        ["i", "j", "d", "q", "b", "s", "dd", "k"],
        [flat_vector(BIGINT, maybe(rng.integers(-50, 50, n).tolist())),
         flat_vector(INTEGER, maybe(rng.integers(-9, 10, n).tolist())),
         flat_vector(DOUBLE, maybe(rng.choice([0.0, -0.0, 1.5, -2.25, NAN, 1e300, 7.0, float("inf")], n).tolist())),
         flat_vector(DOUBLE, maybe(np.round(rng.normal(0, 10, n), 2).tolist())),
         flat_vector(BOOLEAN, maybe((rng.random(n) < 0.5).tolist())),
         dictionary_vector(VARCHAR, rng.integers(0, 6, n), ["PROMO TIN", "STANDARD", "promo", None, "", "PROMOx_"]),
         dictionary_vector(DOUBLE, rng.integers(0, 5, n), [1.5, None, NAN, -4.0, 0.5], index_nulls=rng.random(n) < 0.05),
         constant_vector(BIGINT, 7, n)])


EXPRS = [
    ("+", ("col", "i"), ("lit", T_BIG, 5)),
    ("*", ("col", "q"), ("-", ("lit", T_DBL, 1.0), ("col", "dd"))),
    ("/", ("col", "q"), ("col", "d")),
    ("%", ("col", "i"), ("col", "k")),
    ("neg", ("col", "q")),
    ("<", ("col", "d"), ("col", "q")),
    (">=", ("col", "d"), ("col", "dd")),
    ("=", ("col", "d"), ("col", "d")),
    ("<>", ("col", "i"), ("col", "k")),
    ("between", ("col", "q"), ("lit", T_DBL, -5.0), ("lit", T_DBL, 5.0)),
    ("and", ("col", "b"), (">", ("col", "i"), ("lit", T_BIG, 0))),
    ("or", ("col", "b"), (">", ("col", "i"), ("lit", T_BIG, 0))),
    ("not", ("col", "b")),
    ("is_null", ("col", "dd")),
    ("case", (">", ("col", "i"), ("lit", T_BIG, 0)), ("col", "q"), ("lit", T_DBL, 0.0)),
    ("case", ("col", "b"), ("col", "i")),
    ("cast", T_DBL, ("col", "i")),
    ("cast", T_BIG, ("col", "j")),
    ("cast", T_BIG, ("col", "q")),
    ("cast", T_BOOL, ("col", "q")),
    ("cast", T_BOOL, ("col", "i")),
    ("like", "s", "PROMO%"),
    ("like", "s", "%o_"),
    ("strcmp", "s", "<", "Q"),
    ("case", ("like", "s", "PROMO%"), ("*", ("col", "q"), ("lit", T_DBL, 2.0)), ("lit", T_DBL, 0.0)),
]


def test_generated_project_kernels_match_the_oracle(tmp):
    rv = table()
    got, err, sqls = run_projections(rv, EXPRS, tmp)
    assert err == 0
    plan = PlanBuilder().values(rv.names, rv.types).project([f"{s} as p{i}" for i, s in enumerate(sqls)]).planNode()
    want = pyoracle.run_plan(plan, [rv], threads=1).rows()
    assert len(want) == rv.size
    for c, sql in enumerate(sqls):
        for r in range(rv.size):
            assert same(got[c][r], want[r][c]), (sql, r, got[c][r], want[r][c])


@pytest.mark.parametrize("expr", [
    ("+", ("col", "i"), ("lit", T_BIG, 2**63 - 10)),        # BIGINT overflow (CheckedArithmetic.h:27-60)
    ("/", ("col", "i"), ("-", ("col", "k"), ("lit", T_BIG, 7))),  # division by zero
    ("cast", T_INT, ("*", ("col", "i"), ("lit", T_BIG, 2**31))),  # out of range for INTEGER
    ("cast", T_BIG, ("col", "d")),                           # NaN / 1e300 / inf cannot be cast to BIGINT
])
def test_generated_kernels_raise_where_the_oracle_raises(expr, tmp):
    rng = np.random.default_rng(1)
    n = 64
    rv = row_vector(["i", "d", "k"], [flat_vector(BIGINT, rng.integers(1, 50, n).tolist()),
                                      flat_vector(DOUBLE, rng.choice([1.5, NAN, 1e300, float("inf")], n).tolist()), constant_vector(BIGINT, 7, n)])
    _, err, sqls = run_projections(rv, [expr], tmp)
    assert err != 0, sqls
    with pytest.raises(pyoracle.OracleUserError):
        pyoracle.run_plan(PlanBuilder().values(rv.names, rv.types).project([f"{sqls[0]} as p"]).planNode(), [rv], threads=1)


def run_filter(rv, expr, tmp):
    """The filter form of the generated kernel: one selection bit per row (TRUE and not NULL), as processFilterResults
    (exec/OperatorUtils.cpp:231-321) reads a filter's result."""
    asm = Asm([(n, c.type) for n, c in zip(rv.names, rv.columns)])
    reg, _, sql = asm.emit(expr)
    cols = [c.to_c() for c in rv.columns]
    n = rv.size
    ia, ca = (Instr * len(asm.ins))(*asm.ins), (Const * max(1, len(asm.consts)))(*asm.consts)
    prog = Program(ia, len(asm.ins), len(asm.ins), reg, asm.reg, ca, len(asm.consts), 0)
    L = lib()
    L.vb2k_expression_jit_compiles.restype = C.c_int32
    buf = C.create_string_buffer(1 << 18)
    arr = (CColumn * len(cols))(*cols)
    oa = (Output * 1)()
    rc = L.vb2k_expression_jit_compiles(C.byref(prog), arr, len(cols), 1, oa, 0, buf, len(buf))
    assert rc == 1, buf.value.decode(errors="replace")[:3000]
    K = _host_kernel(buf.value.decode(), tmp)
    a = JitArgs()
    for i, c in enumerate(cols):
        a.cols[i] = c
    for i, c in enumerate(asm.consts):
        a.consts[i] = c
    a.n = n
    bits = np.zeros((n + 31) // 32 + 2, dtype=np.uint32)
    err = np.zeros(2, dtype=np.int32)
    a.sel_bits, a.error_flag = bits.ctypes.data, err.ctypes.data
    K.run_on_host(C.byref(a))
    return [i for i in range(n) if (int(bits[i >> 5]) >> (i & 31)) & 1], int(err[0]), sql


@pytest.mark.parametrize("expr", [
    ("and", ("between", ("col", "q"), ("lit", T_DBL, -5.0), ("lit", T_DBL, 5.0)), ("<", ("col", "i"), ("lit", T_BIG, 24))),
    ("or", ("like", "s", "PROMO%"), ("and", ("col", "b"), (">", ("col", "dd"), ("lit", T_DBL, 0.0)))),
    ("not", ("or", ("is_null", ("col", "d")), ("<", ("col", "d"), ("col", "q")))),
    (">", ("col", "d"), ("lit", T_DBL, 1e300)),   # only NaN and +inf are larger
])
def test_generated_filter_kernels_select_the_oracle_rows(expr, tmp):
    rv = table(n=500, seed=3)
    ids = flat_vector(BIGINT, np.arange(rv.size))
    selected, err, sql = run_filter(rv, expr, tmp)
    assert err == 0
    with_id = row_vector(rv.names + ["id"], rv.columns + [ids])
    plan = PlanBuilder().values(with_id.names, with_id.types).filter(sql).project(["id"]).planNode()
    want = [r[0] for r in pyoracle.run_plan(plan, [with_id], threads=1).rows()]
    assert selected == want, sql


# ---- the real expression compiler in front: SQL -> vb2_plan_expression_program -> generated source -> host -------------
def compiled_node(plan, ordinal=0):
    """The program expr_compiler.cpp produces for the ordinal-th Filter / Project node of a plan."""
    L = lib()
    ins, consts = (Instr * 256)(), (Const * 32)()
    strings = C.create_string_buffer(4096)
    header = (C.c_int32 * 7)()
    regs, types, ident = (C.c_int32 * 32)(), (C.c_int32 * 32)(), (C.c_int32 * 32)()
    err = C.create_string_buffer(2048)
    rc = L.vb2_plan_expression_program(plan.sexpr.encode(), ordinal, ins, 256, consts, 32, strings, 4096, header, regs, types, ident, 32, err, 2048)
    assert rc == 0, err.value.decode()
    return {"ins": ins, "consts": consts, "strings": strings, "n_ins": header[0], "n_consts": header[1], "n_filter": header[2], "filter_reg": header[3],
            "n_regs": header[4], "outs": [(regs[i], types[i], ident[i]) for i in range(header[5])], "is_filter": bool(header[6])}


def run_compiled(rv, node, tmp):
    """Runs a compiled node's program over rv on the host. Filter: list of selected rows. Project: one python list per
    computed output (None entries for identity projections, which never reach the kernel)."""
    cols = [c.to_c() for c in rv.columns]
    n = rv.size
    prog = Program(node["ins"], node["n_ins"], node["n_filter"], node["filter_reg"], node["n_regs"], node["consts"], node["n_consts"], 0)
    outs, bufs = [], []
    for reg, t, _ in node["outs"]:
        if reg < 0 or node["is_filter"]:
            continue
        width = {T_DBL: 8, T_BIG: 8, T_INT: 4, T_BOOL: 1}[t]
        vals = np.zeros(n * width + 8, dtype=np.uint8)
        nulls = np.zeros((n + 31) // 32 + 2, dtype=np.uint32)
        bufs.append((vals, nulls, t))
        outs.append(Output(reg, t, vals.ctypes.data, nulls.ctypes.data))
    L = lib()
    L.vb2k_expression_jit_compiles.restype = C.c_int32
    buf = C.create_string_buffer(1 << 18)
    arr = (CColumn * len(cols))(*cols)
    oa = (Output * max(1, len(outs)))(*outs)
    rc = L.vb2k_expression_jit_compiles(C.byref(prog), arr, len(cols), 1 if node["is_filter"] else 0, oa, len(outs), buf, len(buf))
    assert rc == 1, buf.value.decode(errors="replace")[:3000]
    K = _host_kernel(buf.value.decode(), tmp)
    a = JitArgs()
    for i, c in enumerate(cols):
        a.cols[i] = c
    for i in range(node["n_consts"]):
        a.consts[i] = node["consts"][i]
    for i, o in enumerate(outs):
        a.outs[i] = o
    a.n = n
    bits = np.zeros((n + 31) // 32 + 2, dtype=np.uint32)
    err = np.zeros(2, dtype=np.int32)
    a.sel_bits, a.error_flag = bits.ctypes.data, err.ctypes.data
    K.run_on_host(C.byref(a))
    assert err[0] == 0
    if node["is_filter"]:
        return [i for i in range(n) if (int(bits[i >> 5]) >> (i & 31)) & 1]
    result, it = [], iter(bufs)
    for reg, t, _ in node["outs"]:
        if reg < 0:
            result.append(None)
            continue
        vals, nulls, _ = next(it)
        dt = {T_DBL: np.float64, T_BIG: np.int64, T_INT: np.int32, T_BOOL: np.uint8}[t]
        v = vals[: n * np.dtype(dt).itemsize].view(dt)
        valid = [(int(nulls[i >> 5]) >> (i & 31)) & 1 for i in range(n)]
        result.append([None if not ok else (bool(x) if t == T_BOOL else (float(x) if t == T_DBL else int(x))) for x, ok in zip(v.tolist(), valid)])
    return result


def lineitem(n=400, seed=2, nulls=True):
    from velox_b200 import tpch
    rng = np.random.default_rng(seed)

    def maybe(v, p=0.08):
        return [None if nulls and rng.random() < p else x for x in v]

    return row_vector(
        ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate", "l_partkey", "p_type", "id"],
        [flat_vector(DOUBLE, maybe(rng.integers(1, 51, n).astype(float).tolist())),
         flat_vector(DOUBLE, maybe(np.round(rng.uniform(900, 105000, n), 2).tolist())),
         flat_vector(DOUBLE, maybe((rng.integers(0, 11, n) / 100).tolist())),
         flat_vector(DOUBLE, maybe((rng.integers(0, 9, n) / 100).tolist())),
         flat_vector(INTEGER, maybe(rng.integers(8036, 10562, n).tolist())),
         flat_vector(BIGINT, maybe(rng.integers(1, 2000, n).tolist())),
         dictionary_vector(VARCHAR, rng.integers(0, len(tpch.PTYPE_DICT), n), tpch.PTYPE_DICT),
         flat_vector(BIGINT, np.arange(n))])


@pytest.mark.parametrize("nulls", [False, True])
def test_tpch_expressions_through_the_expression_compiler(nulls, tmp):
    """The projections and filters of TPC-H Q1 / Q6 / Q14 (exec/tests/utils/TpchQueryBuilder.cpp:203-256,756-788,1639-1702):
    SQL -> expression compiler (common sub-expressions shared) -> generated kernel source -> host execution, against the
    oracle; with and without NULLs in the inputs."""
    rv = lineitem(nulls=nulls)
    projections = ["l_extendedprice * (1.0 - l_discount) AS disc_price", "l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS charge",
                   "l_extendedprice * l_discount AS q6", "(CASE WHEN (p_type LIKE 'PROMO%') THEN l_extendedprice * (1.0 - l_discount) ELSE 0.0 END) AS promo",
                   "100.00 * l_discount / l_tax AS ratio", "l_quantity", "cast(l_partkey as double) + l_quantity AS mixed"]
    plan = PlanBuilder().values(rv.names, rv.types).project(projections).planNode()
    node = compiled_node(plan)
    got = run_compiled(rv, node, tmp)
    want = pyoracle.run_plan(plan, [rv], threads=1).rows()
    assert [g is None for g in got] == [False, False, False, False, False, True, False]  # l_quantity is an identity projection
    for c, col in enumerate(got):
        if col is None:
            continue
        for r in range(rv.size):
            assert same(col[r], want[r][c]), (projections[c], r, col[r], want[r][c])
    filters = ["l_shipdate < '1998-09-03'::DATE",
               "l_shipdate between '1994-01-01'::DATE and '1994-12-31'::DATE and l_discount between 0.05 and 0.07 and l_quantity < 24.0",
               "l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE",
               "(l_quantity < 10.0 or p_type like '%BRASS') and not (l_tax is null) and l_partkey % 3 = 1"]
    for f in filters:
        fplan = PlanBuilder().values(rv.names, rv.types).filter(f).project(["id"]).planNode()
        fnode = compiled_node(fplan, 0)
        assert fnode["is_filter"]
        selected = run_compiled(rv, fnode, tmp)
        assert selected == [r[0] for r in pyoracle.run_plan(fplan, [rv], threads=1).rows()], f


# ---- randomised differential run: random typed SQL expressions through compiler + generated source vs the oracle ---------
def _rand_expr(rng, typ, depth):
    """A random SQL expression of type 'i' (BIGINT), 'd' (DOUBLE) or 'b' (BOOLEAN) over the columns of table(). Integer
    arithmetic stays far from overflow and never divides by a column (error semantics inside AND / OR / CASE are allowed to
    differ between a left-to-right evaluator and one that evaluates every branch)."""
    leaf = depth <= 0 or rng.random() < 0.25
    if typ == "i":
        if leaf:
            return rng.choice(["i", "k", "cast(j as bigint)", str(rng.randint(-9, 9))])
        form = rng.choice(["+", "-", "*", "%", "neg", "case", "cast_d", "cast_b"])
        if form in "+-*":
            return f"({_rand_expr(rng, 'i', depth - 1)} {form} {_rand_expr(rng, 'i', depth - 2)})"
        if form == "%":
            return f"({_rand_expr(rng, 'i', depth - 1)} % {rng.choice([3, 7, -5])})"
        if form == "neg":
            return f"(-{_rand_expr(rng, 'i', depth - 1)})"
        if form == "case":
            return f"(case when {_rand_expr(rng, 'b', depth - 1)} then {_rand_expr(rng, 'i', depth - 1)} else {_rand_expr(rng, 'i', depth - 2)} end)"
        if form == "cast_d":
            return "cast(q as bigint)"
        return f"cast({_rand_expr(rng, 'b', depth - 1)} as bigint)"
    if typ == "d":
        if leaf:
            return rng.choice(["d", "q", "dd", repr(rng.choice([0.0, 1.5, -2.25, 0.05, 100.0]))])
        form = rng.choice(["+", "-", "*", "/", "neg", "case", "case_noelse", "cast"])
        if form in "+-*/":
            return f"({_rand_expr(rng, 'd', depth - 1)} {form} {_rand_expr(rng, 'd', depth - 2)})"
        if form == "neg":
            return f"(-{_rand_expr(rng, 'd', depth - 1)})"
        if form == "case":
            return f"(case when {_rand_expr(rng, 'b', depth - 1)} then {_rand_expr(rng, 'd', depth - 1)} else {_rand_expr(rng, 'd', depth - 2)} end)"
        if form == "case_noelse":
            return f"(case when {_rand_expr(rng, 'b', depth - 1)} then {_rand_expr(rng, 'd', depth - 1)} end)"
        return f"cast({_rand_expr(rng, 'i', depth - 1)} as double)"
    if leaf:
        return rng.choice(["b", "(i > 0)", "(d < q)", "(s like 'PROMO%')", "(dd is null)", "(j = 3)"])
    form = rng.choice(["cmp_i", "cmp_d", "and", "or", "not", "between", "is_null", "like", "cast"])
    if form == "cmp_i":
        return f"({_rand_expr(rng, 'i', depth - 1)} {rng.choice(['<', '<=', '>', '>=', '=', '<>'])} {_rand_expr(rng, 'i', depth - 2)})"
    if form == "cmp_d":
        return f"({_rand_expr(rng, 'd', depth - 1)} {rng.choice(['<', '<=', '>', '>=', '=', '<>'])} {_rand_expr(rng, 'd', depth - 2)})"
    if form in ("and", "or"):
        return f"({_rand_expr(rng, 'b', depth - 1)} {form} {_rand_expr(rng, 'b', depth - 1)})"
    if form == "not":
        return f"(not {_rand_expr(rng, 'b', depth - 1)})"
    if form == "between":
        return f"({_rand_expr(rng, 'd', depth - 1)} between {_rand_expr(rng, 'd', depth - 2)} and {_rand_expr(rng, 'd', depth - 2)})"
    if form == "is_null":
        return f"({_rand_expr(rng, rng.choice('id'), depth - 1)} is null)"
    if form == "like":
        return f"(s like '{rng.choice(['PROMO%', '%o', '_ROMO%', '%', 'STANDARD', '%x_'])}')"
    return f"cast({_rand_expr(rng, 'i', depth - 1)} as boolean)"


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_expressions_through_compiler_and_generated_source(seed, tmp):
    import random
    rng = random.Random(seed)
    rv = table(n=256, seed=seed)
    checked = 0
    for _ in range(10):  # a program holds at most 64 registers / 32 constants: a few expressions per plan
        exprs = [_rand_expr(rng, rng.choice("idb"), 3) for _ in range(3)]
        plan = PlanBuilder().values(rv.names, rv.types).project([f"{e} as p{i}" for i, e in enumerate(exprs)]).planNode()
        want = pyoracle.run_plan(plan, [rv], threads=1).rows()
        try:
            node = compiled_node(plan)
        except AssertionError as ex:
            if "more than" in str(ex):  # program limits (registers / constants / instructions): the operator raises the same way
                continue
            raise
        got = run_compiled(rv, node, tmp)
        for c, col in enumerate(got):
            if col is None:  # the expression folded to a plain column reference
                continue
            checked += 1
            for r in range(rv.size):
                assert same(col[r], want[r][c]), (exprs[c], r, col[r], want[r][c])
    assert checked >= 15


# ---- the interpreter (expr_vm.cu), the JIT's fallback and parity partner, on the host ------------------------------------
VM_SHIM = r"""
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>
#include "velox_b200_kernels.h"
// Not `vb2`: the product library's host-side launch stubs of these kernel templates carry the same mangled names, and a
// weak template symbol binds to whichever definition the process loaded first.
namespace vb2_on_host {
struct Dim3 { unsigned x = 0, y = 0, z = 0; };
static Dim3 threadIdx, blockIdx, gridDim;
static int phase = 0;
static size_t ballot_at = 0;
static std::vector<unsigned> masks;
static inline unsigned __ballot_sync(unsigned, bool p) {
  const size_t i = ballot_at++;
  if (phase == 0) {
    if (masks.size() <= i) masks.resize(i + 1, 0u);
    if (p) masks[i] |= 1u << (threadIdx.x & 31u);
    return 0u;
  }
  return masks[i];
}
static inline int atomicCAS(int* p, int expect, int value) { const int old = *p; if (old == expect) *p = value; return old; }
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __ddiv_rn(double a, double b) { return a / b; }
static inline int64_t __mul64hi(int64_t a, int64_t b) { return static_cast<int64_t>((static_cast<__int128>(a) * b) >> 64); }
using std::isnan;
using std::round;
using std::fmod;
#define __global__
#define __device__
#define __forceinline__ inline
#define __grid_constant__
#define __launch_bounds__(x)
#define __shared__
#define __restrict__
#include "vm_ops.inc"
uint64_t vm_regs[64 * 256];  // the interpreter's register file: [reg][thread] in shared memory on the device
// ---- expr_vm.cu, from its limits to the end of the project kernel ----
%s
}  // namespace vb2_on_host
extern "C" void run_vm_on_host(const vb2_program* prog, const vb2_column* cols, int ncols, const vb2_output* outs, int nouts, long long n,
                               unsigned* sel_bits, int* error_flag, int filter, int plain) {
  using namespace vb2_on_host;
  static VmArgs a;
  std::memset(&a, 0, sizeof(a));
  std::memcpy(a.instrs, prog->instrs, sizeof(vb2_instr) * prog->n_instrs);
  std::memcpy(a.consts, prog->consts, sizeof(vb2_const) * prog->n_consts);
  std::memcpy(a.cols, cols, sizeof(vb2_column) * ncols);
  if (nouts) std::memcpy(a.outs, outs, sizeof(vb2_output) * nouts);
  a.n_instrs = prog->n_instrs;
  a.n_filter_instrs = prog->n_filter_instrs;
  a.filter_reg = prog->filter_reg;
  a.n_outs = nouts;
  a.n = n;
  a.sel_bits = sel_bits;
  a.error_flag = error_flag;
  gridDim.x = 1;
  blockIdx.x = 0;
  for (unsigned warp = 0; warp < 8; ++warp) {
    masks.clear();
    for (phase = 0; phase < 2; ++phase)
      for (unsigned lane = 0; lane < 32; ++lane) {
        threadIdx.x = warp * 32 + lane;
        ballot_at = 0;
        if (filter) { if (plain) vm_filter_kernel<true>(a); else vm_filter_kernel<false>(a); }
        else { if (plain) vm_project_kernel<true>(a); else vm_project_kernel<false>(a); }
      }
  }
}
"""


@pytest.fixture(scope="module")
def vm(tmp_path_factory):
    text = open(os.path.join(ROOT, "velox_b200", "csrc", "expr_vm.cu")).read()
    begin = text.index("constexpr int kVmMaxInstrs")
    end = text.index("// ---- selection bitmap -> ascending row numbers")
    d = tmp_path_factory.mktemp("vm_on_host")
    src = d / "vm.cpp"
    src.write_text(VM_SHIM % text[begin:end])
    so = d / "libvm.so"
    subprocess.check_call(["g++", "-std=c++20", "-O1", "-fPIC", "-shared", "-Wl,-Bsymbolic", "-ffp-contract=off", "-w", "-I", os.path.join(ROOT, "include"),
                           "-I", os.path.join(ROOT, "velox_b200", "csrc"), "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.run_vm_on_host.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_longlong, C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return L


def run_interpreter(vm, rv, node):
    cols = [c.to_c() for c in rv.columns]
    n = rv.size
    prog = Program(node["ins"], node["n_ins"], node["n_filter"], node["filter_reg"], node["n_regs"], node["consts"], node["n_consts"], 0)
    outs, bufs = [], []
    for reg, t, _ in node["outs"]:
        if reg < 0 or node["is_filter"]:
            continue
        vals = np.zeros(n * 8 + 8, dtype=np.uint8)  # the interpreter stores 8-byte words except INTEGER / BOOLEAN
        nulls = np.zeros((n + 31) // 32 + 2, dtype=np.uint32)
        bufs.append((vals, nulls, t))
        outs.append(Output(reg, t, vals.ctypes.data, nulls.ctypes.data))
    arr = (CColumn * len(cols))(*cols)
    oa = (Output * max(1, len(outs)))(*outs)
    bits = np.zeros((n + 31) // 32 + 2, dtype=np.uint32)
    err = np.zeros(2, dtype=np.int32)
    vm.run_vm_on_host(C.byref(prog), arr, len(cols), oa, len(outs), n, bits.ctypes.data, err.ctypes.data, 1 if node["is_filter"] else 0, 0)
    assert err[0] == 0
    if node["is_filter"]:
        return [i for i in range(n) if (int(bits[i >> 5]) >> (i & 31)) & 1]
    result, it = [], iter(bufs)
    for reg, t, _ in node["outs"]:
        if reg < 0:
            result.append(None)
            continue
        vals, nulls, _ = next(it)
        dt = {T_DBL: np.float64, T_BIG: np.int64, T_INT: np.int32, T_BOOL: np.uint8}[t]
        v = vals[: n * np.dtype(dt).itemsize].view(dt)
        valid = [(int(nulls[i >> 5]) >> (i & 31)) & 1 for i in range(n)]
        result.append([None if not ok else (bool(x) if t == T_BOOL else (float(x) if t == T_DBL else int(x))) for x, ok in zip(v.tolist(), valid)])
    return result


@pytest.mark.parametrize("seed", [11, 12])
def test_interpreter_on_host_matches_the_oracle_and_the_jit(seed, vm, tmp):
    """The interpreter kernels of expr_vm.cu compiled for the host run the compiler's programs for the TPC-H expressions and
    for random expressions: equal to the oracle, and to what the JIT's generated source computes, value for value."""
    import random
    rng = random.Random(seed)
    rv = table(n=200, seed=seed)
    for _ in range(8):
        exprs = [_rand_expr(rng, rng.choice("idb"), 3) for _ in range(3)]
        plan = PlanBuilder().values(rv.names, rv.types).project([f"{e} as p{i}" for i, e in enumerate(exprs)]).planNode()
        try:
            node = compiled_node(plan)
        except AssertionError as ex:
            if "more than" in str(ex):
                continue
            raise
        want = pyoracle.run_plan(plan, [rv], threads=1).rows()
        got_vm, got_jit = run_interpreter(vm, rv, node), run_compiled(rv, node, tmp)
        for c, col in enumerate(got_vm):
            if col is None:
                continue
            for r in range(rv.size):
                assert same(col[r], want[r][c]) and same(col[r], got_jit[c][r]), (exprs[c], r, col[r], want[r][c], got_jit[c][r])
    li = lineitem(n=300, seed=seed)
    f = "l_shipdate between '1994-01-01'::DATE and '1994-12-31'::DATE and l_discount between 0.05 and 0.07 and l_quantity < 24.0"
    fplan = PlanBuilder().values(li.names, li.types).filter(f).project(["id"]).planNode()
    assert run_interpreter(vm, li, compiled_node(fplan)) == [r[0] for r in pyoracle.run_plan(fplan, [li], threads=1).rows()]
