"""ORDER BY on the device (csrc/sort.cu: the rank sort of small inputs and the stable LSD radix sort — per-block digit
histograms, one scan, warp-ranked stable scatter — of large ones) compiled FOR THE HOST and run under the lock-step
emulation of tests/host_emulator.py, against Python's stable sort under the reference's order (core::SortOrder,
velox/core/PlanNode.h:64-95: NULLs first or last independent of the direction; NaN largest and -0 = +0,
velox/type/FloatingPointUtil.h:52-98; ties keep input order). No GPU needed."""
import ctypes as C
import functools
import math

import numpy as np
import pytest

from host_emulator import between, build, source

BIGINT, INTEGER, DOUBLE, BOOLEAN = 4, 3, 6, 0

BODY = r"""
// ---- sort.cu: key codes, rank sort, radix passes ----
%(sort)s
}  // namespace vb2_on_host
using namespace vb2_on_host;
extern "C" {
void h_rank_sort(const vb2_sort_key* keys, int nkeys, int32_t n, int32_t* order) {
  SortKeys sk{};
  sk.n = nkeys;
  for (int k = 0; k < nkeys; ++k) sk.k[k] = keys[k];
  std::vector<uint64_t> codes(static_cast<size_t>(n) * nkeys);
  std::vector<uint8_t> ranks(static_cast<size_t>(n) * nkeys);
  const unsigned blocks = (n + kThreads - 1) / kThreads;
  launch(blocks, kThreads, [&] { sort_encode_all_kernel(sk, n, codes.data(), ranks.data()); });
  switch (nkeys) {
    case 1: launch(blocks, kThreads, [&] { rank_sort_kernel<1>(codes.data(), ranks.data(), n, nkeys, order); }); break;
    case 2: launch(blocks, kThreads, [&] { rank_sort_kernel<2>(codes.data(), ranks.data(), n, nkeys, order); }); break;
    default: launch(blocks, kThreads, [&] { rank_sort_kernel<3>(codes.data(), ranks.data(), n, nkeys, order); });
  }
}
// the radix path of vb2k_sort_order, launch for launch
void h_radix_sort(const vb2_sort_key* keys, int nkeys, int64_t n, int32_t* order) {
  const int64_t B = (n + 4095) / 4096 > 0 ? (n + 4095) / 4096 : 1;
  const int64_t tile = ((n + B - 1) / B + kThreads - 1) / kThreads * kThreads;
  std::vector<uint64_t> ka(n), kb(n);
  std::vector<int32_t> vb(n);
  std::vector<uint32_t> hist(B * 256);
  const unsigned g = static_cast<unsigned>((n + kThreads - 1) / kThreads);
  int32_t* vin = order;
  int32_t* vout = vb.data();
  launch(g, kThreads, [&] { iota_kernel(vin, n); });
  auto pass = [&](uint64_t*& kin, uint64_t*& kout, int shift) {
    launch(static_cast<unsigned>(B), kThreads, [&] { radix_hist_kernel(kin, n, tile, shift, hist.data()); });
    launch(1, 1024, [&] { scan_u32_kernel(hist.data(), B * 256); });
    launch(static_cast<unsigned>(B), kThreads, [&] { radix_scatter_kernel(kin, vin, kout, vout, n, tile, shift, hist.data()); });
    std::swap(kin, kout);
    std::swap(vin, vout);
  };
  for (int k = nkeys - 1; k >= 0; --k) {
    const vb2_sort_key& key = keys[k];
    uint64_t *kin = ka.data(), *kout = kb.data();
    launch(g, kThreads, [&] { sort_encode_key_kernel(key, vin, n, kin); });
    const int bits = key_width(key);
    for (int shift = 0; shift < bits; shift += 8) pass(kin, kout, shift);
    if (key.nulls) {
      launch(g, kThreads, [&] { sort_null_rank_kernel(key, vin, n, kin); });
      pass(kin, kout, 0);
    }
  }
  if (vin != order) std::memcpy(order, vin, static_cast<size_t>(n) * 4);
}
}
"""


class SortKey(C.Structure):
    _fields_ = [("values", C.c_void_p), ("nulls", C.c_void_p), ("type", C.c_int32), ("ascending", C.c_int32), ("nulls_first", C.c_int32),
                ("significant_bits", C.c_int32)]


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    body = BODY % {"sort": between(source("sort.cu"), "constexpr int kThreads = 256;", "inline unsigned grid_for")}
    return build(tmp_path_factory.mktemp("sort_on_host"), "sort", body)


def _validity(values):
    n = len(values)
    words = np.zeros((n + 63) // 64 + 1, dtype=np.uint64)
    for i, v in enumerate(values):
        if v is not None:
            words[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
    return words


def _reference_order(columns, specs):
    """Stable order of the row numbers under the reference's comparison. specs: (ascending, nulls_first) per key."""
    def cmp_values(a, b):
        if isinstance(a, float) or isinstance(b, float):
            an, bn = math.isnan(a), math.isnan(b)
            if an or bn:
                return 0 if an and bn else (1 if an else -1)  # NaN is the largest value
        return (a > b) - (a < b)

    def cmp_rows(i, j):
        for col, (asc, nulls_first) in zip(columns, specs):
            a, b = col[i], col[j]
            if a is None or b is None:
                if a is None and b is None:
                    continue
                return (-1 if nulls_first else 1) if a is None else (1 if nulls_first else -1)
            c = cmp_values(a, b)
            if c:
                return c if asc else -c
        return 0

    return sorted(range(len(columns[0])), key=functools.cmp_to_key(cmp_rows))


def _keys(columns, types, specs, keep):
    arr = (SortKey * len(columns))()
    for k, (col, typ, (asc, nf)) in enumerate(zip(columns, types, specs)):
        if typ == BOOLEAN:
            bits = np.zeros((len(col) + 63) // 64 + 1, dtype=np.uint64)
            for i, v in enumerate(col):
                if v:
                    bits[i >> 6] |= np.uint64(1) << np.uint64(i & 63)
            vals = bits
        else:
            dt = {BIGINT: np.int64, INTEGER: np.int32, DOUBLE: np.float64}[typ]
            vals = np.array([0 if v is None else v for v in col], dtype=dt)
        nulls = _validity(col) if any(v is None for v in col) else None
        keep += [vals, nulls]
        arr[k] = SortKey(vals.ctypes.data, nulls.ctypes.data if nulls is not None else None, typ, int(asc), int(nf), 0)
    return arr


def _table(n, seed):
    rng = np.random.default_rng(seed)
    nan = float("nan")
    d = [None if rng.random() < 0.08 else float(v) for v in rng.choice([0.0, -0.0, nan, 1.5, -2.25, 1e300, -1e300, float("inf")], n)]
    i = [None if rng.random() < 0.08 else int(v) for v in rng.integers(-5, 6, n)]
    j = [int(v) for v in rng.integers(-2**31, 2**31 - 1, n)]
    b = [bool(v) for v in rng.random(n) < 0.5]
    return d, i, j, b


@pytest.mark.parametrize("specs", [((True, True), (False, False), (True, True)), ((False, False), (True, True), (False, True))])
def test_rank_sort(host, specs):
    d, i, j, b = _table(700, 1)
    columns, types = [i, d, b], [BIGINT, DOUBLE, BOOLEAN]
    keep = []
    keys = _keys(columns, types, specs, keep)
    order = np.full(len(i), -1, dtype=np.int32)
    host.h_rank_sort(keys, 3, len(i), order.ctypes.data_as(C.c_void_p))
    assert order.tolist() == _reference_order(columns, specs)


def test_radix_sort(host):
    """Two keys — DOUBLE descending with NULLs first (NaN / -0 / infinities), then INTEGER dictionary rank codes with an
    8-bit hint (one pass) ascending with NULLs last: LSD passes compose because every pass is stable. (One case: each of the
    eleven passes runs a 1024-thread scan under the emulation.)"""
    d, i, _, _ = _table(1500, 2)
    code = [None if v is None else v + 5 for v in i]
    columns, types, specs = [d, code], [DOUBLE, INTEGER], ((False, True), (True, False))
    keep = []
    keys = _keys(columns, types, specs, keep)
    keys[1].significant_bits = 8
    order = np.full(len(d), -1, dtype=np.int32)
    host.h_radix_sort(keys, 2, C.c_int64(len(d)), order.ctypes.data_as(C.c_void_p))
    assert order.tolist() == _reference_order(columns, specs)
