"""The radix partitioner in front of the L2-sliced high-cardinality aggregation (csrc/radix_partition.cu: per-chunk
histograms over the 256 partitions + a HyperLogLog sketch, the three-kernel exclusive scan, the shared-memory staged
scatter of keys and payload columns) compiled FOR THE HOST and run under the lock-step emulation of
tests/host_emulator.py, launched in the order vb2k_radix_histogram / vb2k_radix_scatter launch them. A row's partition
is the top byte of folly::hasher's twang_mix64 over its normalized key (the bits a hash-mode group table places it by,
exec/HashTable.cpp:485-519 is the reference's CPU answer to the same cache problem); checked against the oracle's
twang_mix64: partition starts, every partition's rows (keys and payloads still paired), and the sketch registers bit
for bit. No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from host_emulator import between, build, source
from oracle import pyoracle

BODY = r"""
static inline int __ffsll(long long v) { return __builtin_ffsll(v); }
static inline int atomicMax(int* p, int v) {
  int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
  while (v > old && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
  return old;
}
// ---- common.cuh: hash mixers ----
%(mixers)s
// ---- radix_partition.cu ----
%(radix)s
}  // namespace vb2_on_host
using namespace vb2_on_host;
extern "C" {
int h_hll_registers() { return 1 << kHllBits; }
int h_chunk_rows() { return kRadixChunkRows; }
void h_radix(const uint64_t* norm, const void* values, int is64, int64_t min, int64_t rows, int grid, const void* const* cols, void* const* cols_out,
             const int32_t* col_bytes, int ncols, uint64_t* keys_out, int64_t* part_start, int32_t* hll) {
  const int64_t nchunks = (rows + kRadixChunkRows - 1) / kRadixChunkRows;
  std::vector<int32_t> hist(nchunks * kRadixParts, -1);
  std::vector<int64_t> base(nchunks * kRadixParts, -1), seg_sums(kOffsetSegs * kRadixParts, -1);
  RadixKey k{norm, values, is64, min};
  RadixCols c{};
  c.n = ncols;
  for (int i = 0; i < ncols; ++i) { c.in[i] = cols[i]; c.out[i] = cols_out[i]; c.bytes[i] = col_bytes[i]; }
  std::memset(hll, 0, sizeof(int32_t) << kHllBits);
  launch(grid, kRadixThreads, [&] { radix_hist_kernel(k, rows, nchunks, hist.data(), hll); });
  launch(kOffsetSegs, kRadixParts, [&] { radix_offsets_a_kernel(hist.data(), nchunks, seg_sums.data()); });
  launch(1, kRadixParts, [&] { radix_offsets_b_kernel(seg_sums.data(), part_start); });
  launch(kOffsetSegs, kRadixParts, [&] { radix_offsets_c_kernel(hist.data(), nchunks, seg_sums.data(), base.data()); });
  launch(grid, kRadixThreads, [&] { radix_scatter_kernel(k, rows, nchunks, base.data(), keys_out, c); });
}
}
"""


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    common, radix = source("common.cuh"), source("radix_partition.cu")
    body = BODY % {
        "mixers": between(common, "__host__ __device__ __forceinline__ uint64_t twang_mix64", "__device__ __forceinline__ uint64_t hash_f64"),
        "radix": between(radix, "constexpr int kRadixParts", "}  // namespace vb2"),
    }
    assert "asm" not in body
    return build(tmp_path_factory.mktemp("radix_on_host"), "radix", body)


A = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


def _mix(norm):
    L = pyoracle.lib()
    return np.array([L.orc_twang_mix64(int(k)) for k in norm], dtype=np.uint64)


def _check(host, norm, key_args, grid):
    n = len(norm)
    rowid = np.arange(n, dtype=np.int64)
    tag = (np.arange(n) * 3 + 1).astype(np.int32)
    out_rowid, out_tag = np.full(n, -1, dtype=np.int64), np.full(n, -1, dtype=np.int32)
    cols = (C.c_void_p * 2)(rowid.ctypes.data, tag.ctypes.data)
    outs = (C.c_void_p * 2)(out_rowid.ctypes.data, out_tag.ctypes.data)
    widths = np.array([8, 4], dtype=np.int32)
    keys_out = np.zeros(n, dtype=np.uint64)
    part_start = np.full(257, -1, dtype=np.int64)
    hll = np.full(host.h_hll_registers(), -1, dtype=np.int32)
    host.h_radix(*key_args, C.c_int64(n), grid, cols, outs, A(widths), 2, A(keys_out), A(part_start), A(hll))
    hashes = _mix(norm)
    pid = (hashes >> np.uint64(56)).astype(np.int64)
    assert np.array_equal(part_start, np.concatenate([[0], np.cumsum(np.bincount(pid, minlength=256))]))
    # a permutation of the rows, grouped by partition, keys and both payloads still paired
    assert np.array_equal(np.sort(out_rowid), rowid)
    assert np.array_equal(keys_out, norm[out_rowid]) and np.array_equal(out_tag, tag[out_rowid])
    assert np.array_equal(pid[out_rowid], np.sort(pid))
    # inside a partition the chunks follow one another (chunk c's run of the partition lies before chunk c + 1's)
    chunk = out_rowid // host.h_chunk_rows()
    for p in np.unique(pid)[:16]:
        run = chunk[part_start[p]:part_start[p + 1]]
        assert (np.diff(run) >= 0).all()
    # the sketch: 1/8 of the rows (low three hash bits zero), 4096 registers indexed by the next 12 bits, rank = position of
    # the first set bit of what lies between those and the partition byte, a stop bit bounding it
    want = np.zeros(len(hll), dtype=np.int32)
    for h in hashes[(hashes & np.uint64(7)) == 0].tolist():
        idx = (h >> 3) & 4095
        rest = (h >> 15) | (1 << 41)
        want[idx] = max(want[idx], (rest & -rest).bit_length())
    assert np.array_equal(hll, want) and (want > 0).sum() > 100
    return part_start


def test_normalized_keys_three_chunks_and_a_partial_one(host):
    rng = np.random.default_rng(5)
    n = 3 * 8192 + 100
    norm = rng.integers(1, 1 << 40, n).astype(np.uint64)
    norm[1000:1400] = norm[7]  # a hot key: 400 rows of one chunk land in one partition
    _check(host, norm, (A(norm), None, 0, C.c_int64(0)), grid=2)


@pytest.mark.parametrize("is64", [0, 1])
def test_flat_integer_key_column(host, is64):
    """No normalisation pass for one flat integer key: normalized key = value - min + 1 computed on the fly
    (vb2k_normalize_keys with one column, VectorHasher value ids: exec/VectorHasher.cpp:560-640)."""
    rng = np.random.default_rng(6 + is64)
    n = 8192 + 4000
    lo = -5000
    values = rng.integers(lo, 3_000_000, n).astype(np.int64 if is64 else np.int32)
    norm = (values.astype(np.int64) - lo + 1).astype(np.uint64)
    _check(host, norm, (None, A(values), is64, C.c_int64(lo)), grid=3)  # more blocks than chunks: the third one idles


def test_fewer_rows_than_one_chunk(host):
    norm = np.arange(1, 778, dtype=np.uint64)
    start = _check_small(host, norm)
    assert start[256] == 777


def _check_small(host, norm):
    n = len(norm)
    keys_out = np.zeros(n, dtype=np.uint64)
    part_start = np.full(257, -1, dtype=np.int64)
    hll = np.full(host.h_hll_registers(), -1, dtype=np.int32)
    host.h_radix(A(norm), None, 0, C.c_int64(0), C.c_int64(n), 1, None, None, None, 0, A(keys_out), A(part_start), A(hll))
    pid = (_mix(norm) >> np.uint64(56)).astype(np.int64)
    assert np.array_equal(part_start, np.concatenate([[0], np.cumsum(np.bincount(pid, minlength=256))]))
    assert np.array_equal(np.sort(keys_out), np.sort(norm))
    assert np.array_equal((_mix(keys_out) >> np.uint64(56)).astype(np.int64), np.sort(pid))
    return part_start
