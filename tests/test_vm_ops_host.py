"""The row-level scalar semantics every kernel shares (velox_b200/csrc/vm_ops.inc: included by common.cuh for the
interpreter and the fused kernels, handed to NVRTC as the prelude of JIT-compiled expressions) compiled FOR THE HOST
and checked on the CPU: against the reference's known-answer vectors (tests/golden/scalar_vectors.json), against the
oracle's scalar kernels on random operands, and LIKE / string comparison against Python. No GPU needed — the same
source text runs on the device."""
import ctypes as C
import json
import os
import random
import re
import subprocess

import pytest

from oracle import pyoracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r"""
#include <cmath>
#include <cstdint>
#define __device__
#define __forceinline__ inline
using std::isnan;
static inline int64_t __mul64hi(int64_t a, int64_t b) { return static_cast<int64_t>((static_cast<__int128>(a) * b) >> 64); }
#include "vm_ops.inc"
extern "C" {
int vo_cmp_f64(int op, double a, double b) { return cmp_f64(op, a, b); }
int vo_cmp_i64(int op, int64_t a, int64_t b) { return cmp_int<int64_t>(op, a, b); }
int vo_fast_f64(int which, double a, double b) {
  switch (which) { case 0: return lt_f64(a, b); case 1: return lte_f64(a, b); case 2: return gt_f64(a, b); case 3: return gte_f64(a, b); default: return eq_f64(a, b); }
}
int vo_checked(int op, int64_t a, int64_t b, int64_t* out) {
  switch (op) { case 0: return add_overflow_i64(a, b, out); case 1: return sub_overflow_i64(a, b, out); default: return mul_overflow_i64(a, b, out); }
}
int vo_like(const char* s, int sl, const char* p, int pl) { return like_match(s, sl, p, pl); }
int vo_strcmp(const char* a, int al, const char* b, int bl) { return str_compare(a, al, b, bl); }
}
"""


@pytest.fixture(scope="module")
def vo(tmp_path_factory):
    d = tmp_path_factory.mktemp("vmops")
    src = d / "harness.cpp"
    src.write_text(HARNESS)
    lib = d / "libvmops.so"
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-I", os.path.join(ROOT, "velox_b200", "csrc"),
                           "-o", str(lib), str(src)])
    L = C.CDLL(str(lib))
    L.vo_cmp_f64.argtypes = [C.c_int, C.c_double, C.c_double]
    L.vo_cmp_i64.argtypes = [C.c_int, C.c_int64, C.c_int64]
    L.vo_fast_f64.argtypes = [C.c_int, C.c_double, C.c_double]
    L.vo_checked.argtypes = [C.c_int, C.c_int64, C.c_int64, C.POINTER(C.c_int64)]
    L.vo_like.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    L.vo_strcmp.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int]
    return L


def _f(x):
    return float(x) if isinstance(x, str) else x


def test_nan_ordering_known_answers(vo):
    """ComparisonsTest.cpp:650-720 / FloatingPointUtil.h:52-98: NaN is the largest value and equals itself."""
    with open(os.path.join(ROOT, "tests", "golden", "scalar_vectors.json")) as f:
        g = json.load(f)["nan_ordering"]
    for case in g["cases"]:
        a, b = _f(case["a"]), _f(case["b"])
        got = [bool(vo.vo_cmp_f64(op, a, b)) for op in (2, 3, 0, 1)]  # gt gte lt lte
        assert got == case["expected"], case
        assert [bool(vo.vo_fast_f64(w, a, b)) for w in (2, 3, 0, 1)] == case["expected"], case
    nan = float("nan")
    assert vo.vo_cmp_f64(4, nan, nan) == 1 and vo.vo_cmp_f64(5, nan, nan) == 0 and vo.vo_fast_f64(4, nan, nan) == 1


def test_comparisons_and_checked_arithmetic_match_the_oracle(vo):
    L = pyoracle.lib()
    rng = random.Random(7)
    specials = [0.0, -0.0, float("nan"), float("inf"), -float("inf"), 1.5, -1.5, 5e-324, 1.7976931348623157e308]
    doubles = specials + [rng.uniform(-1e6, 1e6) for _ in range(200)]
    for _ in range(4000):
        a, b = rng.choice(doubles), rng.choice(doubles)
        for op in range(6):
            assert bool(vo.vo_cmp_f64(op, a, b)) == bool(L.orc_compare_f64(op, a, b)), (op, a, b)
    edges = [0, 1, -1, 2**31, -2**31, 2**32, 2**62, -2**62, 2**63 - 1, -2**63, 3037000499, 3037000500, -3037000500]
    ints = edges + [rng.randint(-2**63, 2**63 - 1) for _ in range(300)] + [rng.randint(-2**33, 2**33) for _ in range(300)]
    out_v, out_o = C.c_int64(), C.c_int64()
    for _ in range(6000):
        a, b = rng.choice(ints), rng.choice(ints)
        for op in range(3):
            ov = vo.vo_checked(op, a, b, C.byref(out_v)) != 0
            oo = L.orc_checked_i64(op, a, b, C.byref(out_o)) != 0
            exact = (a + b, a - b, a * b)[op]
            assert ov == oo == (not -2**63 <= exact <= 2**63 - 1), (op, a, b)
            if not ov:
                assert out_v.value == out_o.value == exact
        for op in range(6):
            want = (a < b, a <= b, a > b, a >= b, a == b, a != b)[op]
            assert bool(vo.vo_cmp_i64(op, a, b)) == want


def test_multiply_overflow_known_answer(vo):
    """ArithmeticTest.cpp:236-240: the smallest integer times -1 overflows (checked arithmetic, CheckedArithmetic.h:27-60)."""
    out = C.c_int64()
    assert vo.vo_checked(2, -2**63, -1, C.byref(out)) != 0
    assert vo.vo_checked(2, -2**62, 2, C.byref(out)) == 0 and out.value == -2**63
    assert vo.vo_checked(0, 2**63 - 1, 1, C.byref(out)) != 0 and vo.vo_checked(1, -2**63, 1, C.byref(out)) != 0


def _like_regex(pattern: bytes):
    parts = []
    for ch in pattern.decode("latin-1"):
        parts.append(".*" if ch == "%" else "." if ch == "_" else re.escape(ch))
    return re.compile("^" + "".join(parts) + "$", re.S)


def test_like_and_string_compare_match_python(vo):
    """LIKE with % and _ (no escape; functions/lib/Re2Functions.cpp:710-733 is the prefix fast path of these semantics) and
    bytewise string comparison (StringView::compare)."""
    rng = random.Random(11)
    alphabet = "abPROM%_ "
    for _ in range(20000):
        s = "".join(rng.choice("abPROM ") for _ in range(rng.randint(0, 9))).encode()
        p = "".join(rng.choice(alphabet) for _ in range(rng.randint(0, 6))).encode()
        assert bool(vo.vo_like(s, len(s), p, len(p))) == bool(_like_regex(p).match(s.decode("latin-1"))), (s, p)
    assert vo.vo_like(b"PROMO BRUSHED TIN", 17, b"PROMO%", 6) == 1 and vo.vo_like(b"STANDARD PROMO", 14, b"PROMO%", 6) == 0
    for _ in range(5000):
        a = bytes(rng.randint(0, 255) for _ in range(rng.randint(0, 5)))
        b = bytes(rng.randint(0, 255) for _ in range(rng.randint(0, 5)))
        want = -1 if a < b else (1 if a > b else 0)
        assert vo.vo_strcmp(a, len(a), b, len(b)) == want, (a, b)
