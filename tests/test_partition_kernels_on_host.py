"""The stable partition order of B200PartitionedOutput (csrc/hash_partition.cu: per-block histograms, the parallel offsets
scan, the warp-ranked stable scatter — HashPartitionFunction::partition followed by the per-destination grouping of
exec/PartitionedOutput.cpp) the selection-bitmap expansion behind every filter (csrc/expr_vm.cu sel_count / scan /
write: processFilterResults' selectedIndices, exec/OperatorUtils.cpp:231-321) and the exclusive scan that turns per-row match
counts into the join's output offsets (csrc/hash_join.cu) compiled FOR THE HOST and run under the
lock-step emulation of tests/host_emulator.py (an OS thread per CUDA thread, barriers for __syncthreads and the warp
collectives). Results against numpy: the stable argsort by partition id — ids given, or computed on the fly from a BIGINT
key as folly::hasher (twang_mix64) % partitions, bit-exact with the oracle's routing — and the ascending row numbers of
the set bits. No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from host_emulator import between, build, source
from oracle import pyoracle

BODY = r"""
// ---- common.cuh: hash mixers, warp reductions ----
%(mixers)s
%(reductions)s
// ---- hash_partition.cu: the stable partition order ----
%(part)s
// ---- expr_vm.cu: selection bitmap -> ascending row numbers ----
%(sel)s
// ---- hash_join.cu: exclusive scan of the per-row match counts ----
%(scan)s
}  // namespace vb2_on_host
using namespace vb2_on_host;
extern "C" {
void h_partition_order(const uint32_t* ids, const void* key, int is64, int64_t rows, int parts, int64_t* counts, int32_t* order) {
  const PartSrc src{ids, key, is64};
  const int64_t nblocks = (rows + kPartRowsPerBlock - 1) / kPartRowsPerBlock;
  std::vector<int32_t> hist(nblocks * parts);
  std::vector<int64_t> base(nblocks * parts);
  launch(static_cast<unsigned>(nblocks), kPartThreads, [&] { part_hist_kernel(src, rows, parts, hist.data()); });
  launch(1, kOffsetThreads, [&] { part_offsets_kernel(hist.data(), nblocks, parts, counts, base.data()); });
  launch(static_cast<unsigned>(nblocks), kPartThreads, [&] { part_scatter_kernel(src, rows, parts, base.data(), order); });
}
void h_bits_to_indices(const uint32_t* bits, int64_t rows, int32_t* indices, int64_t* count) {
  const int64_t nwords = (rows + 31) >> 5;
  const int64_t nblocks = (nwords + kSelWordsPerBlock - 1) / kSelWordsPerBlock;
  std::vector<int32_t> counts(nblocks);
  std::vector<int64_t> offsets(nblocks);
  launch(static_cast<unsigned>(nblocks), kSelThreads, [&] { sel_count_kernel(bits, nwords, counts.data()); });
  launch(1, 1024, [&] { sel_scan_kernel(counts.data(), nblocks, offsets.data(), count); });
  launch(static_cast<unsigned>(nblocks), kSelThreads, [&] { sel_write_kernel(bits, nwords, offsets.data(), indices); });
}
void h_exclusive_scan(const int32_t* in, int64_t n, int64_t* out, int64_t* total) {
  const int64_t nblocks = (n + kScanThreads * kScanItems - 1) / (kScanThreads * kScanItems);
  std::vector<int64_t> sums(nblocks);
  launch(static_cast<unsigned>(nblocks), kScanThreads, [&] { scan_block_sums_kernel(in, n, sums.data()); });
  launch(1, 1024, [&] { scan_offsets_kernel(sums.data(), nblocks, total); });
  launch(static_cast<unsigned>(nblocks), kScanThreads, [&] { scan_write_kernel(in, n, sums.data(), out); });
}
}
"""


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    common, part, vm, join = source("common.cuh"), source("hash_partition.cu"), source("expr_vm.cu"), source("hash_join.cu")
    body = BODY % {
        "mixers": between(common, "__host__ __device__ __forceinline__ uint64_t twang_mix64", "__device__ __forceinline__ uint64_t hash_f64"),
        "reductions": between(common, "__device__ __forceinline__ double warp_sum(double v)", "}  // namespace vb2"),
        "part": between(part, "constexpr int kPartThreads", "// --- fixed-capacity segments"),
        "sel": between(vm, "constexpr int kSelThreads", "static unsigned vm_grid"),
        "scan": between(join, "constexpr int kScanThreads", "static unsigned grid_for"),
    }
    return build(tmp_path_factory.mktemp("partition_on_host"), "part", body)


P = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731


@pytest.mark.parametrize("parts", [2, 7, 64])
def test_stable_partition_order_with_given_ids(host, parts):
    rng = np.random.default_rng(parts)
    n = 3 * 4096 + 1234  # four blocks, the last one partial
    ids = rng.integers(0, parts, n).astype(np.uint32)
    ids[100:3000] = 1 % parts  # a long run of one partition
    counts = np.zeros(parts, dtype=np.int64)
    order = np.full(n, -1, dtype=np.int32)
    host.h_partition_order(P(ids), None, 0, C.c_int64(n), parts, P(counts), P(order))
    assert np.array_equal(counts, np.bincount(ids, minlength=parts))
    assert np.array_equal(order, np.argsort(ids, kind="stable").astype(np.int32))


def test_partition_ids_from_a_key_column_match_the_oracle(host):
    """part_id computed on the fly = HashPartitionFunction over one BIGINT key: twang_mix64(key) % partitions
    (exec/HashPartitionFunction.cpp:75-118, exec/VectorHasher.cpp:62-126), the routing a CPU worker would compute."""
    rng = np.random.default_rng(1)
    n, parts = 2 * 4096 + 77, 8
    keys = rng.integers(-2**62, 2**62, n)
    counts = np.zeros(parts, dtype=np.int64)
    order = np.full(n, -1, dtype=np.int32)
    host.h_partition_order(None, P(keys), 1, C.c_int64(n), parts, P(counts), P(order))
    L = pyoracle.lib()
    want_ids = np.array([L.orc_twang_mix64(int(k) & 0xFFFFFFFFFFFFFFFF) % parts for k in keys], dtype=np.int64)
    assert np.array_equal(counts, np.bincount(want_ids, minlength=parts))
    assert np.array_equal(order, np.argsort(want_ids, kind="stable").astype(np.int32))


@pytest.mark.parametrize("density", [0.5, 0.02, 0.0, 1.0])
def test_selection_bitmap_expands_to_ascending_rows(host, density):
    """Dense bitmaps take the warp-per-word path, sparse ones the per-thread bit loop; several blocks of 32 768 rows."""
    rng = np.random.default_rng(int(density * 100))
    n = 2 * 32768 + 4321
    keep = rng.random(n) < density
    words = np.zeros((n + 31) // 32 + 2, dtype=np.uint32)
    for i in np.nonzero(keep)[0]:
        words[i >> 5] |= np.uint32(1) << np.uint32(i & 31)
    indices = np.full(n + 1, -1, dtype=np.int32)
    count = np.zeros(1, dtype=np.int64)
    host.h_bits_to_indices(P(words), C.c_int64(n), P(indices), P(count))
    want = np.nonzero(keep)[0].astype(np.int32)
    assert int(count[0]) == len(want)
    assert np.array_equal(indices[:len(want)], want)


def test_exclusive_scan_of_match_counts(host):
    """count -> scan -> emit: the offsets of the join's output pairs (listJoinResults, exec/HashTable.cpp:2133-2350)."""
    rng = np.random.default_rng(4)
    n = 3 * 2048 + 555
    counts = rng.integers(0, 5, n).astype(np.int32)
    out = np.full(n, -1, dtype=np.int64)
    total = np.zeros(1, dtype=np.int64)
    host.h_exclusive_scan(P(counts), C.c_int64(n), P(out), P(total))
    want = np.concatenate([[0], np.cumsum(counts.astype(np.int64))])
    assert int(total[0]) == int(want[-1]) and np.array_equal(out, want[:-1])
