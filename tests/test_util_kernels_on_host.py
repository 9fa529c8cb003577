"""The small bitmap / index kernels every operator leans on (csrc/util.cu: selection of validity bits through row
numbers, bits <-> bytes for the exchange, counts -> validity for aggregate results, index validity for outer joins,
AND of bitmaps, widening / narrowing, scatter) and the join slot flags of the fused probe (csrc/fused_scan.cu) compiled
FOR THE HOST and run under the lock-step emulation of tests/host_emulator.py, against numpy. Sizes are not multiples of
32 or 64 and the grids have both more and fewer warps than words: the partial last word and the grid-stride loops are
what can go wrong here. Velox keeps validity as 64-bit words with bit i of word i / 64 = row i
(velox/common/base/BitUtil.h); the 32-bit words the ballots write are the same bytes on a little-endian machine.
No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from host_emulator import between, build, source

BODY = r"""
// ---- util.cu ----
%(util)s
// ---- fused_scan.cu: join slot flags ----
%(flags)s
}  // namespace vb2_on_host
using namespace vb2_on_host;
extern "C" {
void h_gather_bits(const uint64_t* in, const int32_t* sel, int64_t n, uint32_t* out, int grid, int threads) { launch(grid, threads, [&] { gather_bits_kernel(in, sel, n, out); }); }
void h_unpack_bits(const uint64_t* in, int64_t n, uint8_t* out, int grid, int threads) { launch(grid, threads, [&] { unpack_bits_kernel(in, n, out); }); }
void h_pack_bools(const uint8_t* in, int64_t n, uint32_t* out, int grid, int threads) { launch(grid, threads, [&] { pack_bools_kernel(in, n, out); }); }
void h_positive_bits(const int64_t* counts, int64_t n, uint32_t* out, int grid, int threads) { launch(grid, threads, [&] { positive_bits_kernel(counts, n, out); }); }
void h_index_validity(const int32_t* idx, int64_t n, uint32_t* valid, int32_t* clamped, int grid, int threads) {
  launch(grid, threads, [&] { index_validity_kernel(idx, n, valid, clamped); });
}
void h_and_bits(const uint64_t* a, const uint64_t* b, int64_t nwords, uint64_t* out, int grid, int threads) { launch(grid, threads, [&] { and_bits_kernel(a, b, nwords, out); }); }
void h_widen(const int32_t* in, int64_t n, int64_t* out, int negate, int grid, int threads) { launch(grid, threads, [&] { widen_i32_kernel(in, n, out, negate != 0); }); }
void h_narrow(const int64_t* in, int64_t n, int32_t* out, int grid, int threads) { launch(grid, threads, [&] { narrow_i64_kernel(in, n, out); }); }
void h_scatter64(const int64_t* in, const int32_t* src, const int32_t* dst, int64_t n, int64_t* out, int grid, int threads) {
  launch(grid, threads, [&] { scatter_kernel<int64_t>(in, src, dst, n, out); });
}
void h_fill_iota(uint64_t* a, int32_t* b, int32_t* c, int64_t n, int grid, int threads) {
  launch(grid, threads, [&] { fill_u64_kernel(a, n, 0xfeedfacecafebeefull); });
  launch(grid, threads, [&] { fill_i32_kernel(b, n, -3); });
  launch(grid, threads, [&] { iota_i32_kernel(c, n); });
}
void h_join_slot_flags(const int32_t* head, const int32_t* codes, const uint8_t* flag, int64_t range, uint8_t* out, int grid, int threads) {
  launch(grid, threads, [&] { join_slot_flags_kernel(head, codes, flag, range, out); });
}
}
"""


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    util, fused = source("util.cu"), source("fused_scan.cu")
    body = BODY % {
        "util": between(util, "__global__ void fill_u64_kernel", "}  // namespace vb2"),
        "flags": between(fused, "__global__ void join_slot_flags_kernel", "static std::deque<Entry>& registry()"),
    }
    return build(tmp_path_factory.mktemp("util_on_host"), "util", body)


A = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
SHAPES = [(1, 64), (3, 96), (2, 256)]  # (blocks, threads): 2, 9 and 16 warps


def _bits64(flags):
    """Velox's layout: 64-bit words, bit i of word i / 64 = row i."""
    padded = np.zeros((len(flags) + 63) // 64 * 64, dtype=bool)
    padded[:len(flags)] = flags
    return np.packbits(padded.reshape(-1, 8), axis=1, bitorder="little").ravel().view(np.uint64)


def _words32(flags):
    padded = np.zeros((len(flags) + 31) // 32 * 32, dtype=bool)
    padded[:len(flags)] = flags
    return np.packbits(padded.reshape(-1, 8), axis=1, bitorder="little").ravel().view(np.uint32)


@pytest.mark.parametrize("grid,threads", SHAPES)
@pytest.mark.parametrize("n", [1, 31, 1000, 1057])
def test_bitmaps(host, grid, threads, n):
    rng = np.random.default_rng(n + grid)
    nw = (n + 31) // 32
    # validity of the rows a selection names: out bit k = in bit sel[k]
    src_rows = 3 * n + 70
    flags = rng.random(src_rows) < 0.6
    sel = rng.integers(0, src_rows, n).astype(np.int32)
    out = np.full(nw + 2, 0xABABABAB, dtype=np.uint32)
    host.h_gather_bits(A(_bits64(flags)), A(sel), C.c_int64(n), A(out), grid, threads)
    assert np.array_equal(out[:nw], _words32(flags[sel])) and (out[nw:] == 0xABABABAB).all()
    # bits -> bytes -> bits
    row_flags = rng.random(n) < 0.5
    as_bytes = np.full(n + 3, 9, dtype=np.uint8)
    host.h_unpack_bits(A(_bits64(row_flags)), C.c_int64(n), A(as_bytes), grid, threads)
    assert np.array_equal(as_bytes[:n], row_flags.astype(np.uint8)) and (as_bytes[n:] == 9).all()
    as_bytes[:n] *= rng.integers(1, 200, n).astype(np.uint8)  # any non-zero byte is true
    back = np.full(nw + 2, 0xABABABAB, dtype=np.uint32)
    host.h_pack_bools(A(as_bytes), C.c_int64(n), A(back), grid, threads)
    assert np.array_equal(back[:nw], _words32(row_flags)) and (back[nw:] == 0xABABABAB).all()
    # SUM / MIN / MAX of a group are NULL when no non-null input reached it: validity = count > 0
    counts = rng.integers(-1, 3, n).astype(np.int64)
    out = np.full(nw + 2, 0xABABABAB, dtype=np.uint32)
    host.h_positive_bits(A(counts), C.c_int64(n), A(out), grid, threads)
    assert np.array_equal(out[:nw], _words32(counts > 0)) and (out[nw:] == 0xABABABAB).all()
    # outer-join build side: row -1 = no match -> NULL, index clamped to a readable row
    idx = rng.integers(-1, 50, n).astype(np.int32)
    valid, clamped = np.full(nw + 2, 0xABABABAB, dtype=np.uint32), np.full(n + 2, -9, dtype=np.int32)
    host.h_index_validity(A(idx), C.c_int64(n), A(valid), A(clamped), grid, threads)
    assert np.array_equal(valid[:nw], _words32(idx >= 0)) and (valid[nw:] == 0xABABABAB).all()
    assert np.array_equal(clamped[:n], np.maximum(idx, 0)) and (clamped[n:] == -9).all()
    # AND of two bitmaps; a missing second one is all ones
    a, b = _bits64(rng.random(n) < 0.7), _bits64(rng.random(n) < 0.7)
    out = np.zeros(len(a) + 1, dtype=np.uint64)
    host.h_and_bits(A(a), A(b), C.c_int64(len(a)), A(out), grid, threads)
    assert np.array_equal(out[:-1], a & b) and out[-1] == 0
    host.h_and_bits(A(a), None, C.c_int64(len(a)), A(out), grid, threads)
    assert np.array_equal(out[:-1], a)


@pytest.mark.parametrize("grid,threads", SHAPES)
def test_widen_narrow_scatter_fill(host, grid, threads):
    rng = np.random.default_rng(grid)
    n = 777
    v = rng.integers(-5, 5, n).astype(np.int32)
    wide = np.full(n + 1, 99, dtype=np.int64)
    host.h_widen(A(v), C.c_int64(n), A(wide), 0, grid, threads)
    assert np.array_equal(wide[:n], v) and wide[n] == 99
    host.h_widen(A(v), C.c_int64(n), A(wide), 1, grid, threads)  # negate: 1 where the input is 0 (anti-join match counts)
    assert np.array_equal(wide[:n], (v == 0).astype(np.int64))
    big = rng.integers(-2**31, 2**31, n).astype(np.int64)
    small = np.zeros(n, dtype=np.int32)
    host.h_narrow(A(big), C.c_int64(n), A(small), grid, threads)
    assert np.array_equal(small, big.astype(np.int32))
    # out[dst[i]] = in[src ? src[i] : i]
    dst = rng.permutation(n).astype(np.int32)
    src = rng.integers(0, n, n).astype(np.int32)
    out = np.zeros(n, dtype=np.int64)
    host.h_scatter64(A(big), A(src), A(dst), C.c_int64(n), A(out), grid, threads)
    want = np.zeros(n, dtype=np.int64)
    want[dst] = big[src]
    assert np.array_equal(out, want)
    host.h_scatter64(A(big), None, A(dst), C.c_int64(n), A(out), grid, threads)
    want[dst] = big
    assert np.array_equal(out, want)
    a, b, c = np.zeros(n + 1, dtype=np.uint64), np.zeros(n + 1, dtype=np.int32), np.zeros(n + 1, dtype=np.int32)
    host.h_fill_iota(A(a), A(b), A(c), C.c_int64(n), grid, threads)
    assert (a[:n] == 0xFEEDFACECAFEBEEF).all() and a[n] == 0 and (b[:n] == -3).all() and b[n] == 0
    assert np.array_equal(c[:n], np.arange(n)) and c[n] == 0


@pytest.mark.parametrize("with_codes", [False, True])
def test_join_slot_flags(host, with_codes):
    """One byte per key slot of an array-mode join table for the fused probe: 0 = no build row, 1 = match, 2 = match and
    the build-side predicate holds; the predicate is evaluated once per dictionary entry (codes) or once per build row."""
    rng = np.random.default_rng(3)
    rng_slots, build_rows, ncodes = 500, 300, 20
    head = np.zeros(rng_slots, dtype=np.int32)
    slots = rng.choice(rng_slots, build_rows, replace=False)
    head[slots] = np.arange(build_rows) + 1  # build row + 1
    codes = rng.integers(0, ncodes, build_rows).astype(np.int32)
    flag = (rng.random(ncodes if with_codes else build_rows) < 0.4).astype(np.uint8)
    out = np.full(rng_slots + 1, 7, dtype=np.uint8)
    host.h_join_slot_flags(A(head), A(codes) if with_codes else None, A(flag), C.c_int64(rng_slots), A(out), 2, 128)
    want = np.zeros(rng_slots, dtype=np.uint8)
    rows = head[slots] - 1
    want[slots] = np.where(flag[codes[rows] if with_codes else rows] != 0, 2, 1)
    assert np.array_equal(out[:-1], want) and out[-1] == 7
    host.h_join_slot_flags(A(head), None, None, C.c_int64(rng_slots), A(out), 1, 64)  # no predicate: matches are 1
    assert np.array_equal(out[:-1], (head != 0).astype(np.uint8))
