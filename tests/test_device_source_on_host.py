"""The row-level scalar semantics every kernel shares (velox_b200/csrc/vm_ops.inc: included by common.cuh for the
interpreter and the fused kernels, handed to NVRTC as the prelude of JIT-compiled expressions) compiled FOR THE HOST
and checked on the CPU: against the reference's known-answer vectors (tests/golden/scalar_vectors.json), against the
oracle's scalar kernels on random operands, and LIKE / string comparison against Python. No GPU needed — the same
source text runs on the device. The hashing block of common.cuh (twang_mix64, jenkins_rev_mix32, hashMix, the double
canonicalisation and the CRC32-C based hashBytes in its bitwise device form) gets the same treatment: compiled for the
host and compared bit for bit with the oracle, whose hashBytes uses the SSE4.2 crc32 instruction as the reference does."""
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


HASH_HARNESS = r"""
#include <cmath>
#include <cstdint>
#include <cstring>
#define __host__
#define __device__
#define __forceinline__ inline
using std::isnan;
static inline double __longlong_as_double(long long v) { double d; std::memcpy(&d, &v, 8); return d; }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
%s
extern "C" {
uint64_t dh_twang(uint64_t k) { return twang_mix64(k); }
uint32_t dh_jenkins(uint32_t k) { return jenkins_rev_mix32(k); }
uint64_t dh_mix(uint64_t a, uint64_t b) { return hash_mix(a, b); }
uint64_t dh_f64(double v) { return hash_f64(v); }
uint64_t dh_bytes(uint64_t seed, const uint8_t* p, int32_t n) { return hash_bytes(seed, p, n); }
}
"""


@pytest.fixture(scope="module")
def dh(tmp_path_factory):
    text = open(os.path.join(ROOT, "velox_b200", "csrc", "common.cuh")).read()
    begin = text.index("constexpr uint64_t kNullHash = 1;")
    end = text.index("// Warp / block reductions")
    end = text.rindex("// ----", begin, end)
    d = tmp_path_factory.mktemp("devhash")
    src = d / "harness.cpp"
    src.write_text(HASH_HARNESS % text[begin:end])
    lib = d / "libdevhash.so"
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-Wno-unknown-pragmas", "-o", str(lib), str(src)])
    L = C.CDLL(str(lib))
    L.dh_twang.restype = C.c_uint64
    L.dh_twang.argtypes = [C.c_uint64]
    L.dh_jenkins.restype = C.c_uint32
    L.dh_jenkins.argtypes = [C.c_uint32]
    L.dh_mix.restype = C.c_uint64
    L.dh_mix.argtypes = [C.c_uint64, C.c_uint64]
    L.dh_f64.restype = C.c_uint64
    L.dh_f64.argtypes = [C.c_double]
    L.dh_bytes.restype = C.c_uint64
    L.dh_bytes.argtypes = [C.c_uint64, C.c_char_p, C.c_int32]
    return L


def test_device_hash_source_is_bit_exact_with_the_oracle(dh):
    """VectorHasher's hashes (exec/VectorHasher.cpp:62-126, common/base/BitUtil.h:775-784, BitUtil.cpp:177-230) as the
    device computes them, on the CPU: integers, doubles (all NaNs alike, +0 == -0) and strings of every length class of
    bits::hashBytes (< 8, 8..16, 17..24, > 24 bytes, with and without a tail)."""
    L = pyoracle.lib()
    rng = random.Random(3)
    for k in [0, 1, 2**63, 2**64 - 1, 0x9E3779B97F4A7C15] + [rng.getrandbits(64) for _ in range(3000)]:
        assert dh.dh_twang(k) == L.orc_twang_mix64(k)
        assert dh.dh_jenkins(k & 0xFFFFFFFF) == L.orc_jenkins_rev_mix32(k & 0xFFFFFFFF)
        other = rng.getrandbits(64)
        assert dh.dh_mix(k, other) == L.orc_hash_mix(k, other)
    import struct
    nans = [struct.unpack("<d", struct.pack("<Q", b))[0] for b in (0x7ff8000000000000, 0xfff8000000000000, 0x7ff0000000000001, 0x7fffffffffffffff)]
    for v in [0.0, -0.0, 1.0, -1.0, 5e-324, float("inf"), -float("inf")] + nans + [rng.uniform(-1e9, 1e9) for _ in range(2000)]:
        assert dh.dh_f64(v) == L.orc_hash_f64(v), v
    assert dh.dh_f64(0.0) == dh.dh_f64(-0.0) and len({dh.dh_f64(n) for n in nans}) == 1
    for n in list(range(0, 80)) + [127, 128, 129, 1000]:
        for _ in range(20):
            data = bytes(rng.randint(0, 255) for _ in range(n))
            seed = rng.choice([1, 0, rng.getrandbits(64)])
            assert dh.dh_bytes(seed, data, n) == L.orc_hash_bytes(seed, data, n), (n, seed)


SORT_HARNESS = r"""
#include <cmath>
#include <cstdint>
#include <cstring>
#include "velox_b200_kernels.h"
#define __host__
#define __device__
#define __forceinline__ inline
using std::isnan;
static inline int64_t __mul64hi(int64_t a, int64_t b) { return static_cast<int64_t>((static_cast<__int128>(a) * b) >> 64); }
static inline long long __double_as_longlong(double d) { long long v; std::memcpy(&v, &d, 8); return v; }
#include "vm_ops.inc"
%s
extern "C" {
// order-preserving code + null rank of one value, as the sort kernels compute them
uint64_t sk_code(int32_t type, int32_t ascending, int32_t bits, const void* value, int32_t is_null, int32_t nulls_first, int32_t* rank) {
  uint64_t valid = is_null ? 0 : 1;
  vb2_sort_key k{value, &valid, type, ascending, nulls_first, bits};
  bool nl;
  const uint64_t c = encode_key(k, 0, &nl);
  *rank = null_rank(k, nl);
  return c;
}
}
"""


@pytest.fixture(scope="module")
def sk(tmp_path_factory):
    text = open(os.path.join(ROOT, "velox_b200", "csrc", "sort.cu")).read()
    begin = text.index("// bits of a key's code")
    end = text.index("// ---- small inputs: rank sort")
    d = tmp_path_factory.mktemp("sortkeys")
    src = d / "harness.cpp"
    src.write_text(SORT_HARNESS % text[begin:end])
    lib = d / "libsortkeys.so"
    subprocess.check_call(["g++", "-std=c++20", "-O2", "-fPIC", "-shared", "-I", os.path.join(ROOT, "include"), "-I", os.path.join(ROOT, "velox_b200", "csrc"),
                           "-o", str(lib), str(src)])
    L = C.CDLL(str(lib))
    L.sk_code.restype = C.c_uint64
    L.sk_code.argtypes = [C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_int32)]
    return L


def test_sort_key_codes_preserve_the_reference_order(sk):
    """ORDER BY key codes (core::SortOrder, velox/core/PlanNode.h:64-95; NaN largest and -0 == +0,
    velox/type/FloatingPointUtil.h:52-98): code(a) < code(b) exactly when a sorts before b, for both directions; NULLs rank
    first or last independent of the direction."""
    import math
    BIGINT, INTEGER, DOUBLE = 4, 3, 6
    rng = random.Random(5)

    def code(typ, asc, value, bits=0, null=False, nulls_first=True):
        if typ == DOUBLE:
            buf = C.c_double(value if value is not None else 0.0)
        elif typ == BIGINT:
            buf = C.c_int64(value if value is not None else 0)
        else:
            buf = C.c_int32(value if value is not None else 0)
        rank = C.c_int32()
        c = sk.sk_code(typ, 1 if asc else 0, bits, C.cast(C.byref(buf), C.c_void_p), 1 if null else 0, 1 if nulls_first else 0, C.byref(rank))
        return c, rank.value

    def before_f64(a, b):  # a sorts strictly before b ascending: NaN is the largest value, -0 == +0
        an, bn = math.isnan(a), math.isnan(b)
        if an or bn:
            return (not an) and bn
        return a < b

    doubles = [0.0, -0.0, float("nan"), float("inf"), -float("inf"), 5e-324, -5e-324, 1.5, -1.5, 1e308, -1e308] + [rng.uniform(-1e3, 1e3) for _ in range(300)]
    for _ in range(20000):
        a, b = rng.choice(doubles), rng.choice(doubles)
        for asc in (True, False):
            ca, cb = code(DOUBLE, asc, a)[0], code(DOUBLE, asc, b)[0]
            lo, hi = (a, b) if asc else (b, a)
            assert (ca < cb) == before_f64(lo, hi), (a, b, asc)
            assert (ca == cb) == (not before_f64(a, b) and not before_f64(b, a))
    ints64 = [0, 1, -1, 2**63 - 1, -2**63, 2**32, -2**32] + [rng.randint(-2**63, 2**63 - 1) for _ in range(300)]
    ints32 = [0, 1, -1, 2**31 - 1, -2**31] + [rng.randint(-2**31, 2**31 - 1) for _ in range(300)]
    for typ, pool in ((BIGINT, ints64), (INTEGER, ints32)):
        for _ in range(10000):
            a, b = rng.choice(pool), rng.choice(pool)
            for asc in (True, False):
                ca, cb = code(typ, asc, a)[0], code(typ, asc, b)[0]
                assert (ca < cb) == ((a < b) if asc else (a > b)) and (ca == cb) == (a == b)
    # dictionary rank codes: values promised to lie in [0, 2^bits) are their own code, complemented within the width for DESC
    for _ in range(5000):
        bits = rng.randint(1, 20)
        a, b = rng.randrange(1 << bits), rng.randrange(1 << bits)
        for asc in (True, False):
            ca, cb = code(INTEGER, asc, a, bits)[0], code(INTEGER, asc, b, bits)[0]
            assert ca < (1 << bits) and (ca < cb) == ((a < b) if asc else (a > b))
    for asc in (True, False):
        assert code(BIGINT, asc, None, null=True, nulls_first=True)[1] == 0 and code(BIGINT, asc, None, null=True, nulls_first=False)[1] == 2
        assert code(BIGINT, asc, 7)[1] == 1
