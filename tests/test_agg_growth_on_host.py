"""What happens to a group table when a later batch widens the key ranges or outgrows the capacity (csrc/hash_agg.cu:
occupied_bits -> rekey -> group_move into the new table; the reference rehashes in place, exec/HashTable.cpp:830-905
checkSize / rehash, and re-derives value ids when a VectorHasher's range grows, exec/VectorHasher.cpp:560-640) and how
the partial sums of a fused scan reach an existing array-mode table (merge_partials_kernel), compiled FOR THE HOST and
run under the lock-step emulation of tests/host_emulator.py. Also the per-slot readers of the result path (group_keys,
group_gather, group_valid, group_avg). Expected values are computed in Python from the packing rule
key = sum_k id_k * mult_k, id_k = value - min_k + 1, 0 = NULL. No GPU needed."""
import ctypes as C

import numpy as np
import pytest

from host_emulator import between, build, source

EMPTY = 0xFFFFFFFFFFFFFFFF

BODY = r"""
// ---- common.cuh: hash mixers, warp reductions ----
%(mixers)s
%(reductions)s
// ---- hash_agg.cu: key normalisation constants, accumulator updates, the table view ----
%(norm)s
%(update)s
// ---- hash_agg.cu: occupied slots, keys of slots, rekey, move, per-slot readers ----
%(relayout)s
// ---- hash_agg.cu: merge of fused-scan partials ----
%(merge)s
}  // namespace vb2_on_host
using namespace vb2_on_host;
extern "C" {
int h_type_integer() { return VB2_INTEGER; }
int h_type_bigint() { return VB2_BIGINT; }
void h_occupied_bits(const vb2_group_table* t, uint32_t* bits) { launch(2, 96, [&] { occupied_bits_kernel(*t, bits); }); }
void h_rekey(const vb2_group_table* t, const int32_t* slots, int64_t n, int ncols, const int64_t* old_min, const int64_t* new_min, const uint64_t* old_mult,
             const uint64_t* old_range, const uint64_t* new_mult, const int32_t* old_null_reserved, uint64_t* out) {
  RekeyArgs a{};
  a.n = ncols;
  for (int k = 0; k < ncols; ++k) {
    a.old_min[k] = old_min[k]; a.new_min[k] = new_min[k]; a.old_mult[k] = old_mult[k]; a.old_range[k] = old_range[k]; a.new_mult[k] = new_mult[k];
    a.old_null_reserved[k] = old_null_reserved[k];
  }
  launch(2, 64, [&] { rekey_kernel(*t, slots, n, a, out); });
}
void h_group_move(const vb2_group_table* from, const int32_t* slots, const uint64_t* new_keys, int64_t n, const vb2_group_table* to, int64_t* num_groups,
                  int32_t* error_flag) {
  launch(3, 64, [&] { group_move_kernel(*from, slots, new_keys, n, *to, num_groups, error_flag); });
}
void h_group_keys(const vb2_group_table* t, const int32_t* slots, int64_t n, int64_t min, uint64_t mult, uint64_t range, int null_reserved, int type, void* values,
                  uint32_t* valid) {
  launch(2, 64, [&] { group_keys_kernel(*t, slots, n, min, mult, range, null_reserved, type, values, valid); });
}
void h_readers(const vb2_group_table* t, const int32_t* slots, int64_t n, int word, int count_word, uint64_t* gathered, uint32_t* valid, double* avg) {
  launch(2, 64, [&] { group_gather_kernel(*t, slots, n, word, gathered); });
  launch(2, 64, [&] { group_valid_kernel(*t, slots, n, count_word, valid); });
  launch(2, 64, [&] { group_avg_kernel(*t, slots, n, word, count_word, avg); });
}
void h_merge_partials(const vb2_group_table* t, const double* sums, const int64_t* counts, int ngroups, int nproj, const int32_t* word, const int32_t* proj, int n) {
  MergeArgs m{};
  m.n = n;
  for (int k = 0; k < n; ++k) { m.word[k] = word[k]; m.proj[k] = proj[k]; }
  const int total = ngroups * n;
  launch((total + 127) / 128, 128, [&] { merge_partials_kernel(*t, sums, counts, ngroups, nproj, m); });
}
}
"""


class GroupTable(C.Structure):
    _fields_ = [("rows", C.c_void_p), ("capacity", C.c_int64), ("row_words", C.c_int32), ("hash_mode", C.c_int32)]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    common, agg = source("common.cuh"), source("hash_agg.cu")
    update = between(agg, "__device__ __forceinline__ double input_as_f64", "// ---- keyed hash mode (kHash)")
    update = update.replace("extern __shared__ __align__(16) uint64_t srows[];", "static uint64_t srows[8192];")
    body = BODY % {
        "mixers": between(common, "__host__ __device__ __forceinline__ uint64_t twang_mix64", "__device__ __forceinline__ uint64_t hash_f64"),
        "reductions": between(common, "__device__ __forceinline__ double warp_sum(double v)", "}  // namespace vb2"),
        "norm": between(agg, "constexpr int kMaxNormCols", "__global__ void minmax_kernel"),
        "update": update,
        "relayout": between(agg, "__device__ __forceinline__ bool row_occupied", "// Every output column of an aggregation in one launch"),
        "merge": between(agg, "// Partial results of a fused scan", "static unsigned grid_for(int64_t n, int threads, int per_sm"),
    }
    return build(tmp_path_factory.mktemp("agg_growth_on_host"), "growth", body)


A = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
I64 = lambda v: np.array(v, dtype=np.int64)  # noqa: E731
U64 = lambda v: np.array(v, dtype=np.uint64)  # noqa: E731


def _old_table():
    """Array mode, two nullable key columns: a in [10, 14], b in [0, 3]; id 0 of a column = NULL. Row = (occupied, sum, count)."""
    rng = np.random.default_rng(8)
    old_min, old_range, old_mult = [10, 0], [6, 5], [1, 6]
    row_words, capacity = 3, 30
    rows = np.zeros(capacity * row_words, dtype=np.uint64)
    groups = {}
    for a_val in [None, 10, 12, 14]:
        for b_val in [None, 0, 3]:
            if rng.random() < 0.25:
                continue
            ida = 0 if a_val is None else a_val - 10 + 1
            idb = 0 if b_val is None else b_val + 1
            slot = ida * 1 + idb * 6
            total, count = float(rng.integers(-50, 50)) / 4, int(rng.integers(0, 5))
            rows[slot * row_words:(slot + 1) * row_words] = [1, np.array([total]).view(np.uint64)[0], count]
            groups[(a_val, b_val)] = (slot, total, count)
    assert len(groups) >= 8
    return rows, GroupTable(rows.ctypes.data, capacity, row_words, 0), groups, (old_min, old_range, old_mult)


@pytest.mark.parametrize("to_hash_mode", [0, 1])
def test_relayout_when_the_key_ranges_grow(host, to_hash_mode):
    rows, old, groups, (old_min, old_range, old_mult) = _old_table()
    bits = np.zeros(2, dtype=np.uint32)
    host.h_occupied_bits(C.byref(old), A(bits))
    slots = np.array([s for s in range(old.capacity) if (int(bits[s >> 5]) >> (s & 31)) & 1], dtype=np.int32)
    assert sorted(slots.tolist()) == sorted(s for s, _, _ in groups.values())
    n = len(slots)
    # keys of the occupied slots back to column values (the output path) ...
    a_out, a_valid = np.full(n, -99, dtype=np.int32), np.zeros(1, dtype=np.uint32)
    b_out, b_valid = np.full(n, -99, dtype=np.int64), np.zeros(1, dtype=np.uint32)
    host.h_group_keys(C.byref(old), A(slots), C.c_int64(n), C.c_int64(old_min[0]), C.c_uint64(old_mult[0]), C.c_uint64(old_range[0]), 1, host.h_type_integer(),
                      A(a_out), A(a_valid))
    host.h_group_keys(C.byref(old), A(slots), C.c_int64(n), C.c_int64(old_min[1]), C.c_uint64(old_mult[1]), C.c_uint64(old_range[1]), 1, host.h_type_bigint(),
                      A(b_out), A(b_valid))
    by_slot = {s: k for k, (s, _, _) in groups.items()}
    for i, s in enumerate(slots.tolist()):
        a_val, b_val = by_slot[s]
        assert ((int(a_valid[0]) >> i) & 1) == (a_val is not None) and ((int(b_valid[0]) >> i) & 1) == (b_val is not None)
        assert a_out[i] == (a_val if a_val is not None else 0) and b_out[i] == (b_val if b_val is not None else 0)
    # ... and re-encoded for wider ranges: a in [5, 24], b in [-2, 7]
    new_min, new_mult = [5, -2], [1, 21]
    new_keys = np.zeros(n, dtype=np.uint64)
    host.h_rekey(C.byref(old), A(slots), C.c_int64(n), 2, A(I64(old_min)), A(I64(new_min)), A(U64(old_mult)), A(U64(old_range)), A(U64(new_mult)),
                 A(np.array([1, 1], dtype=np.int32)), A(new_keys))
    want_keys = []
    for s in slots.tolist():
        a_val, b_val = by_slot[s]
        want_keys.append((0 if a_val is None else a_val - 5 + 1) * 1 + (0 if b_val is None else b_val + 2 + 1) * 21)
    assert new_keys.tolist() == want_keys and len(set(want_keys)) == n
    # move every group into the new table: array mode (row = key) or hash mode (open addressing on twang_mix64's top bits)
    capacity = 512 if to_hash_mode else 21 * 11
    new_rows = np.zeros(capacity * 3, dtype=np.uint64)
    if to_hash_mode:
        new_rows[0::3] = EMPTY
    to = GroupTable(new_rows.ctypes.data, capacity, 3, to_hash_mode)
    num_groups, err = np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int32)
    host.h_group_move(C.byref(old), A(slots), A(new_keys), C.c_int64(n), C.byref(to), A(num_groups), A(err))
    assert err[0] == 0 and num_groups[0] == (n if to_hash_mode else 0)  # array mode counts groups from the occupied bits instead
    table = new_rows.reshape(capacity, 3)
    occupied = table[:, 0] != (EMPTY if to_hash_mode else 0)
    assert occupied.sum() == n
    got = {}
    for r in np.nonzero(occupied)[0]:
        key = int(table[r, 0]) if to_hash_mode else int(r)
        got[key] = (table[r, 1:2].view(np.float64)[0], int(table[r, 2]))
    for key, s in zip(want_keys, slots.tolist()):
        _, total, count = groups[by_slot[s]]
        assert got[key] == (total, count)
    # per-slot readers over the new table: raw word, validity = count > 0, AVG = sum / count (0.0 behind a NULL)
    new_slots = np.nonzero(occupied)[0].astype(np.int32)
    gathered, valid, avg = np.zeros(n, dtype=np.uint64), np.zeros(1, dtype=np.uint32), np.zeros(n)
    host.h_readers(C.byref(to), A(new_slots), C.c_int64(n), 1, 2, A(gathered), A(valid), A(avg))
    for i, r in enumerate(new_slots.tolist()):
        total, count = table[r, 1:2].view(np.float64)[0], int(table[r, 2])
        assert gathered[i] == table[r, 1] and ((int(valid[0]) >> i) & 1) == (count > 0)
        assert avg[i] == (total / count if count > 0 else 0.0)


def test_a_full_hash_table_reports_error_100(host):
    rows, old, groups, _ = _old_table()
    slots = np.array(sorted(s for s, _, _ in groups.values()), dtype=np.int32)
    n = len(slots)
    new_rows = np.zeros(4 * 3, dtype=np.uint64)  # four slots for eight or more groups
    new_rows[0::3] = EMPTY
    to = GroupTable(new_rows.ctypes.data, 4, 3, 1)
    num_groups, err = np.zeros(1, dtype=np.int64), np.zeros(1, dtype=np.int32)
    host.h_group_move(C.byref(old), A(slots), A(np.arange(1, n + 1, dtype=np.uint64)), C.c_int64(n), C.byref(to), A(num_groups), A(err))
    assert err[0] == 100 and num_groups[0] <= 4


def test_fused_scan_partials_merge_into_an_existing_table(host):
    """sums[g * nproj + p] / counts[g] of one batch's fused scan are added to the rows of an array-mode table (group g =
    row g) that earlier, generic batches filled; groups the batch did not reach (count 0) are left alone -- their SUM must
    not become 0.0 + (-0.0) or gain a spurious non-null count."""
    ngroups, nproj, row_words = 6, 2, 5  # row = (occupied, sum0, sum1, count, rows of sum1's non-null counter)
    rng = np.random.default_rng(12)
    rows = np.zeros(ngroups * row_words, dtype=np.uint64)
    table = rows.reshape(ngroups, row_words)
    before_sums = rng.integers(-40, 40, (ngroups, 2)) / 8
    before_sums[4, 0] = -0.0
    before_counts = rng.integers(0, 9, ngroups)
    table[:, 0] = 1
    table[:, 1:3] = before_sums.view(np.uint64)
    table[:, 3] = before_counts
    table[:, 4] = before_counts
    t = GroupTable(rows.ctypes.data, ngroups, row_words, 0)
    sums = rng.integers(-40, 40, (ngroups, nproj)) / 8
    counts = np.array([3, 0, 7, 1, 0, 2], dtype=np.int64)
    word = np.array([1, 2, 3, 4], dtype=np.int32)
    proj = np.array([0, 1, -1, -1], dtype=np.int32)
    snapshot = table.copy()
    host.h_merge_partials(C.byref(t), A(sums), A(counts), ngroups, nproj, A(word), A(proj), 4)
    for g in range(ngroups):
        if counts[g] == 0:
            assert np.array_equal(table[g], snapshot[g])  # bit for bit: -0.0 stays -0.0
            continue
        assert np.array_equal(table[g, 1:3].view(np.float64), before_sums[g] + sums[g])
        assert table[g, 3] == before_counts[g] + counts[g] and table[g, 4] == before_counts[g] + counts[g] and table[g, 0] == 1
