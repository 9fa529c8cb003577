"""The fused scan kernels (csrc/fused_scan.cuh: the expression templates, the register and shared-memory accumulators,
the array-mode join probe, the block reduction, fused_scan.cu's finalize step) compiled FOR THE HOST and run under the
lock-step emulation of tests/host_emulator.py with the pipelines the product instantiates ahead of time: TPC-H Q6, Q1
and the Q14 probe side (exec/tests/utils/TpchQueryBuilder.cpp:203-256, 756-788, 1639-1702 -- FilterProject +
HashAggregation (+ HashProbe) collapsed into one pass), the scan -> compact step in front of the exchange and the late
materialisation pair (filter bitmap, gather). Every kernel variant runs:
  * the direct-load kernel;
  * the TMA-staged kernels: producer warp, ring of stages, full / empty mbarriers. Their PTX helpers (mbarrier init /
    expect_tx / arrive / try_wait.parity, cp.async.bulk, bar.sync 1) are the only text replaced, by a host model of the
    same objects: an mbarrier is (pending arrivals, pending transaction bytes, phase), a bulk copy is a memcpy followed by
    complete_tx, the named barrier a barrier over the 256 consumer threads. Ring wrap over both parities, uneven tiles
    per block, tails, staged key columns and the accumulators behind the stages are the device code.
What the model cannot show is the hardware's ordering between an issued ld.shared and the next bulk copy into the same
stage (mbar_arrive_after's data dependence, measured on the GPU) -- that stays with the GPU suite. Sums are compared with
the oracle running the unfused plan (row order there, a fixed tree here: relative 1e-12), counts, bitmaps and compacted
rows exactly. No GPU needed."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from host_emulator import ROOT, between, build, source
from oracle import pyoracle
from velox_b200 import tpch
from velox_b200.plan import PlanBuilder
from velox_b200.vector import BIGINT, DOUBLE, INTEGER, VARCHAR, dictionary_vector, flat_vector, row_vector

BODY = r"""
struct double2 { double x, y; };
struct int2 { int x, y; };
struct longlong2 { long long x, y; };
template <class T> static inline T __ldg(const T* p) { return *p; }
static inline double2 ldg_stream_f64x2(const double* p) { return {p[0], p[1]}; }
static inline longlong2 ldg_stream_i64x2(const int64_t* p) { return {p[0], p[1]}; }
static inline double __dsub_rn(double a, double b) { return a - b; }
static inline double __dmul_rn(double a, double b) { return a * b; }
#define VB2_SIG(...) __VA_ARGS__
// ---- common.cuh: warp reductions ----
%(reductions)s
// ---- fused_scan.cuh: expression templates, Accum, the direct-load kernel ----
%(fx)s
// ---- the mbarrier / bulk-copy helpers of the TMA-staged variant, modelled on the host: an mbarrier is (arrivals
// pending, transaction bytes pending, phase); a phase completes when both reach zero; a bulk copy is a memcpy followed
// by complete_tx; try_wait.parity succeeds once the phase of that parity has completed ----
struct HostBar { uint32_t init = 0, pending = 0; int64_t tx = 0; uint32_t phase = 0; };
static std::mutex bar_mu;
static std::unordered_map<const uint64_t*, HostBar> bars;
static inline void bar_settle(HostBar& b) { if (b.pending == 0 && b.tx == 0) { b.phase ^= 1u; b.pending = b.init; } }
static inline void mbar_init(uint64_t* bar, uint32_t count) { std::lock_guard<std::mutex> g(bar_mu); bars[bar] = HostBar{count, count, 0, 0}; }
static inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { std::lock_guard<std::mutex> g(bar_mu); HostBar& b = bars.at(bar); b.tx += bytes; --b.pending; bar_settle(b); }
static inline void mbar_arrive(uint64_t* bar) { std::lock_guard<std::mutex> g(bar_mu); HostBar& b = bars.at(bar); --b.pending; bar_settle(b); }
static inline void mbar_arrive_after(uint64_t* bar, uint64_t, uint64_t) { mbar_arrive(bar); }
static inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { std::lock_guard<std::mutex> g(bar_mu); return bars.at(bar).phase != parity; }
static inline void mbar_wait(uint64_t* bar, uint32_t parity) { while (!mbar_try_wait(bar, parity)) std::this_thread::yield(); }
static inline void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  std::memcpy(dst, src, bytes);
  std::lock_guard<std::mutex> g(bar_mu);
  HostBar& b = bars.at(bar);
  b.tx -= bytes;
  bar_settle(b);
}
alignas(128) static uint8_t dyn_smem[227 * 1024];
static std::unique_ptr<std::barrier<>> consumer_barrier;
// ---- fused_scan.cuh: the TMA-staged kernel ----
%(tma)s
// ---- fused_scan.cu: finalize, the ahead-of-time pipelines ----
%(finalize)s
%(pipelines)s
%(extra_types)s
}  // namespace fx
}  // namespace vb2_on_host
using namespace vb2_on_host;
using namespace vb2_on_host::fx;

// stages == 0: the direct-load kernel; otherwise the TMA-staged one with that many stages (kMaxG == 0: accumulators in
// shared memory behind the stages). Workspace shape and finalize arguments as in fused_scan.cu's launcher.
template <class P, int kMaxG, class KeyT>
static int run(const KernelArgs& a, int grid, int stages, double* sums, int64_t* counts) {
  const int groups = kMaxG == 0 ? a.ngroups : kMaxG;
  const int kvals = groups * (P::kNP + 1);
  std::vector<double> partials(static_cast<size_t>(grid) * kvals, -1.0);
  if (stages == 0) {
    if constexpr (kMaxG > 0) launch(grid, kThreads, [&] { fused_scan_agg_kernel<P, kMaxG, 2, false, KeyT>(a, partials.data()); });
    else return 2;
  } else {
    const size_t need = static_cast<size_t>(stages) * TileLayout<P, sizeof(KeyT)>::stage_bytes((kMaxG == 0 || kMaxG > 1) ? a.nkeys : 0) +
                        (kMaxG == 0 ? SmemAccum<P>::bytes(a.ngroups, kConsumerThreads) : 0);
    if (need > sizeof(dyn_smem) || stages > kMaxStages) return 3;
    std::memset(dyn_smem, 0xff, sizeof(dyn_smem));
    launch(grid, kTmaThreads, [&] { fused_scan_agg_tma_kernel<P, kMaxG, KeyT>(a, stages, partials.data()); });
  }
  launch(kvals, 32, [&] { fused_finalize_kernel(partials.data(), grid, kvals, P::kNP, groups, a.ngroups, sums, counts); });
  return 0;
}

extern "C" {
// which: 6 = Q6, 1 = Q1 (int32 keys, 8 accumulator groups), 14 = Q14 probe side, 2 = sum(b * (c - d)) where a < x by a BIGINT key
int h_fused(int which, const vb2_fused_args* in, int grid, int stages, int smem_acc, double* sums, int64_t* counts) {
  KernelArgs a{};
  for (int c = 0; c < kMaxCols; ++c) a.cols[c] = in->cols[c];
  for (int k = 0; k < VB2_FUSED_MAX_PARAMS; ++k) { a.consts.pf[k] = in->pf[k]; a.consts.pl[k] = in->pl[k]; a.consts.pi[k] = in->pi[k]; }
  a.rows = in->rows;
  a.nkeys = in->nkeys;
  a.ngroups = in->ngroups;
  for (int k = 0; k < VB2_FUSED_MAX_KEYS; ++k) {
    a.key[k] = in->key[k]; a.key_is64[k] = in->key_is64[k]; a.key_mult[k] = in->key_mult[k]; a.key_min[k] = in->key_min[k]; a.key_lut[k] = in->key_lut[k];
  }
  a.join_slot_flags = reinterpret_cast<const uint8_t*>(in->join_slot_flags);
  a.join_min = in->join_min;
  a.join_range = in->join_range;
  a.release_guard = kReleaseGuard;
  switch (which) {
    case 6: return run<Q6, 1, int32_t>(a, grid, stages, sums, counts);
    case 1: return smem_acc ? run<Q1, 0, int32_t>(a, grid, stages, sums, counts) : run<Q1, 8, int32_t>(a, grid, stages, sums, counts);
    case 14: return run<Q14, 1, int32_t>(a, grid, stages, sums, counts);
    case 2: return smem_acc ? run<SumUnderLt, 0, int64_t>(a, grid, stages, sums, counts) : run<SumUnderLt, 4, int64_t>(a, grid, stages, sums, counts);
  }
  return 1;
}
// Q14ScanCompact: rows in the date range -> (l_partkey, revenue) written densely, one reservation per tile
int h_compact(const vb2_fused_args* in, int grid, int stages, void* out_key, void* out_rev, int64_t capacity, unsigned long long* count, int32_t* error_flag) {
  KernelArgs a{};
  for (int c = 0; c < kMaxCols; ++c) a.cols[c] = in->cols[c];
  for (int k = 0; k < VB2_FUSED_MAX_PARAMS; ++k) { a.consts.pf[k] = in->pf[k]; a.consts.pl[k] = in->pl[k]; a.consts.pi[k] = in->pi[k]; }
  a.rows = in->rows;
  a.release_guard = kReleaseGuard;
  CompactArgs o{};
  o.outs[0] = out_key;
  o.outs[1] = out_rev;
  o.capacity = capacity;
  o.count = count;
  o.error_flag = error_flag;
  std::memset(dyn_smem, 0xff, sizeof(dyn_smem));
  consumer_barrier = std::make_unique<std::barrier<>>(kConsumerThreads);
  launch(grid, kTmaThreads, [&] { fused_scan_compact_tma_kernel<Q14ScanCompact>(a, o, stages); });
  return 0;
}
// Late materialisation, Q14: the filter's column alone through the ring -> selection bitmap (+ kept / seen counters);
// then probe + project + aggregate over the surviving row numbers
int h_filter_bits(const vb2_fused_args* in, int grid, int stages, int tile_stride, uint32_t* bits, unsigned long long* counters) {
  KernelArgs a{};
  for (int c = 0; c < kMaxCols; ++c) a.cols[c] = in->cols[c];
  for (int k = 0; k < VB2_FUSED_MAX_PARAMS; ++k) { a.consts.pf[k] = in->pf[k]; a.consts.pl[k] = in->pl[k]; a.consts.pi[k] = in->pi[k]; }
  a.rows = in->rows;
  a.release_guard = kReleaseGuard;
  std::memset(dyn_smem, 0xff, sizeof(dyn_smem));
  launch(grid, kTmaThreads, [&] {
    fused_filter_bits_tma_kernel<typename Q14::FilterView, filter_tile_rows_for(Q14::F::fmask, Q14::F::imask, Q14::F::lmask)>(a, stages, tile_stride, bits, counters);
  });
  return filter_tile_rows_for(Q14::F::fmask, Q14::F::imask, Q14::F::lmask);
}
int h_gather(const vb2_fused_args* in, int grid, const int32_t* sel, const int64_t* nsel, double* sums, int64_t* counts) {
  KernelArgs a{};
  for (int c = 0; c < kMaxCols; ++c) a.cols[c] = in->cols[c];
  for (int k = 0; k < VB2_FUSED_MAX_PARAMS; ++k) { a.consts.pf[k] = in->pf[k]; a.consts.pl[k] = in->pl[k]; a.consts.pi[k] = in->pi[k]; }
  a.rows = in->rows;
  a.ngroups = 1;
  a.join_slot_flags = reinterpret_cast<const uint8_t*>(in->join_slot_flags);
  a.join_min = in->join_min;
  a.join_range = in->join_range;
  constexpr int kvals = Q14::kNP + 1;
  std::vector<double> partials(static_cast<size_t>(grid) * kvals, -1.0);
  launch(grid, kThreads, [&] { fused_gather_agg_kernel<Q14, 1, int32_t>(a, sel, nsel, partials.data()); });
  launch(kvals, 32, [&] { fused_finalize_kernel(partials.data(), grid, kvals, Q14::kNP, 1, 1, sums, counts); });
  return 0;
}
%(extra_drivers)s
const char* h_sig(int which) {
  static std::string s;
  s = which == 6 ? Q6::sig() : which == 1 ? Q1::sig() : Q14::sig();
  return s.c_str();
}
}
"""


def _build(tmpdir, extra_types="", extra_drivers=""):
    common, cuh, cu = source("common.cuh"), source("fused_scan.cuh"), source("fused_scan.cu")
    fx = between(cuh, "namespace fx {", "// TMA-staged variant (main path)")
    # the one PTX statement of the slice: a predicated DADD
    fx, n = re.subn(r'asm\("\{ \.reg \.pred q; setp\.ne\.s32 q, %1, 0; @q add\.rn\.f64 %0, %0, %2; \}"[^;]*;',
                    "if (hit) sum[g][p] = __dadd_rn(sum[g][p], v[p]);", fx)
    assert n == 1 and "asm" not in fx
    tma = between(cuh, "constexpr int kTileRows = 1024;", "#ifndef __CUDACC_RTC__\n// Folds per-block partials")
    helpers = between(tma, "__device__ __forceinline__ uint32_t smem_u32", "// Byte offsets of the referenced columns")
    assert helpers.count("asm volatile") == 6  # init, expect_tx, arrive, arrive_after, try_wait, bulk_load: all modelled above
    tma = tma.replace(helpers, "").replace("::vb2::fx::", "::vb2_on_host::fx::")
    tma, n1 = re.subn(r"extern __shared__ __align__\(128\) uint8_t tile_smem\[\];", "uint8_t* tile_smem = dyn_smem;", tma)
    tma, n2 = re.subn(r'asm volatile\("fence\.mbarrier_init\.release\.cluster;" ::: "memory"\);', "", tma)
    # bar.sync 1, 256: the named barrier of the eight consumer warps (the producer warp never joins it)
    tma, n3 = re.subn(r'asm volatile\("bar\.sync 1, %0;" ::"n"\(kConsumerThreads\) : "memory"\);', "consumer_barrier->arrive_and_wait();", tma)
    assert n1 == 3 and n2 == 3 and n3 == 3 and "asm" not in tma
    body = BODY % {
        "reductions": between(common, "__device__ __forceinline__ double warp_sum(double v)", "}  // namespace vb2"),
        "fx": fx,
        "tma": tma,
        "extra_types": extra_types,
        "extra_drivers": extra_drivers,
        "finalize": between(cu, "__global__ void fused_finalize_kernel", "// join_slot_flags[slot]"),
        "pipelines": between(cu, "using Q6 = ", "static std::once_flag"),
    }
    L = build(tmpdir, "fused", body)
    L.h_sig.restype = C.c_char_p
    return L


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    return _build(tmp_path_factory.mktemp("fused_on_host"))


class FusedArgs(C.Structure):
    """vb2_fused_args (include/velox_b200_kernels.h)."""
    _fields_ = [("cols", C.c_void_p * 8), ("pf", C.c_double * 12), ("pl", C.c_int64 * 12), ("pi", C.c_int32 * 12), ("rows", C.c_int64),
                ("nkeys", C.c_int32), ("ngroups", C.c_int32), ("key", C.c_void_p * 2), ("key_is64", C.c_int32 * 2), ("key_mult", C.c_int32 * 2),
                ("key_min", C.c_int64 * 2), ("key_lut", C.c_void_p * 2), ("join_slot_flags", C.c_void_p), ("join_min", C.c_int64), ("join_range", C.c_int64)]


def test_args_struct_matches_the_header():
    """The ctypes mirror above is only as good as its field list: same order and types as the C declaration."""
    with open(os.path.join(ROOT, "include", "velox_b200_kernels.h")) as f:
        decl = between(f.read(), "typedef struct vb2_fused_args {", "} vb2_fused_args").split("{", 1)[1]
    names = [re.sub(r"\[\w+\]", "", d).strip(" *") for line in re.findall(r"^\s*(?:const\s+)?\w+\*?\s+([^;/{]+);", decl, re.M) for d in line.split(",")]
    assert names == [n for n, _ in FusedArgs._fields_], names


def _lineitem(n, seed, nparts=500):
    return {k: np.ascontiguousarray(v.numpy()) for k, v in tpch.gen_lineitem(n, nparts, seed=seed, device="cpu").items()}


def _vec(h, name):
    if name == "l_returnflag":
        return dictionary_vector(VARCHAR, h[name], tpch.RETURNFLAG_DICT)
    if name == "l_linestatus":
        return dictionary_vector(VARCHAR, h[name], tpch.LINESTATUS_DICT)
    if name == "l_shipdate":
        return flat_vector(INTEGER, h[name])
    if name == "l_partkey":
        return flat_vector(BIGINT, h[name])
    return flat_vector(DOUBLE, h[name])


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p).value


def _arg(a):
    return a.ctypes.data_as(C.c_void_p)


def _run(host, which, args, grid, ngroups, nproj, stages=0, smem_acc=0):
    sums = np.zeros(max(ngroups, 1) * nproj)
    counts = np.zeros(max(ngroups, 1), dtype=np.int64)
    assert host.h_fused(which, C.byref(args), grid, stages, smem_acc, sums.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p)) == 0
    return sums.reshape(-1, nproj), counts


def _close(got, want, n):
    return abs(got - want) <= max(1e-12, n * 2.0 ** -53) * max(1.0, abs(want))


# direct: 517 rows over 4 x 256 threads = the main loop never runs, tail only. TMA-staged: five tiles per block through a
# ring of two stages (every stage refilled twice: both mbarrier parities), a 300-row tail by direct loads; seven tiles
# over two blocks through three stages (4 + 3 tiles); fewer rows than one tile (the ring is never used).
@pytest.mark.parametrize("n,grid,stages", [(3000, 3, 0), (517, 4, 0), (2 * 5 * 1024 + 300, 2, 2), (7 * 1024 + 1, 2, 3), (700, 2, 4)])
def test_q6(host, n, grid, stages):
    assert host.h_sig(6) == b"F:and(between(i0,pi0,pi1),between(f1,pf0,pf1),lt(f2,pf2));P:multiply(f3,f1)"
    h = _lineitem(n, seed=6)
    a = FusedArgs()
    for c, name in enumerate(["l_shipdate", "l_discount", "l_quantity", "l_extendedprice"]):  # first-use order
        a.cols[c] = _ptr(h[name])
    a.pi[0], a.pi[1] = tpch.Q6_SHIP_LO, tpch.Q6_SHIP_HI
    a.pf[0], a.pf[1], a.pf[2] = 0.05, 0.07, 24.0
    a.rows, a.nkeys, a.ngroups = n, 0, 1
    sums, counts = _run(host, 6, a, grid, 1, 1, stages=stages)
    names = ["l_shipdate", "l_extendedprice", "l_quantity", "l_discount"]
    rv = row_vector(names, [_vec(h, c) for c in names])
    plan = (PlanBuilder().values(rv.names, rv.types)
            .filter("l_shipdate between '1994-01-01'::DATE and '1994-12-31'::DATE and l_discount between 0.05 and 0.07 and l_quantity < 24.0")
            .project(["l_extendedprice * l_discount"]).singleAggregation([], ["sum(p0)", "count(0)"]).planNode())
    want = pyoracle.run_plan(plan, [rv])
    ((want_sum, want_count),) = want.rows()
    assert int(counts[0]) == want_count and counts[0] > 0
    assert _close(sums[0, 0], want_sum, n)


# direct-load kernel with eight register groups; TMA-staged kernel with the accumulators in shared memory behind the
# stages (what the launcher picks for 5+ groups when a ring fits), key columns staged with the tile
@pytest.mark.parametrize("n,grid,stages,smem_acc", [(4000, 2, 0, 0), (6 * 1024 + 77, 2, 2, 1)])
def test_q1(host, n, grid, stages, smem_acc):
    assert host.h_sig(1).startswith(b"F:lt(i0,pi0);P:f1|f2|multiply(f2,minus(pf0,f3))|")
    h = _lineitem(n, seed=1)
    a = FusedArgs()
    for c, name in enumerate(["l_shipdate", "l_quantity", "l_extendedprice", "l_discount", "l_tax"]):
        a.cols[c] = _ptr(h[name])
    a.pi[0] = tpch.Q1_SHIPDATE_LT
    a.pf[0] = a.pf[1] = a.pf[2] = 1.0
    rf, ls = h["l_returnflag"].astype(np.int32), h["l_linestatus"].astype(np.int32)
    a.rows, a.nkeys, a.ngroups = n, 2, 6
    a.key[0], a.key[1] = _ptr(rf), _ptr(ls)
    a.key_mult[0], a.key_mult[1] = 2, 1  # gid = returnflag * 2 + linestatus
    sums, counts = _run(host, 1, a, grid, 6 if smem_acc else 8, 5, stages=stages, smem_acc=smem_acc)
    if smem_acc:
        sums, counts = np.vstack([sums, np.zeros((2, 5))]), np.concatenate([counts, [0, 0]])
    names = ["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_shipdate"]
    rv = row_vector(names, [_vec(h, c) for c in names])
    plan = (PlanBuilder().values(rv.names, rv.types).filter("l_shipdate < '1998-09-03'::DATE")
            .project(["l_returnflag", "l_linestatus", "l_quantity", "l_extendedprice", "l_extendedprice * (1.0 - l_discount) AS dp",
                      "l_extendedprice * (1.0 - l_discount) * (1.0 + l_tax) AS ch", "l_discount"])
            .singleAggregation(["l_returnflag", "l_linestatus"], ["sum(l_quantity)", "sum(l_extendedprice)", "sum(dp)", "sum(ch)", "sum(l_discount)", "count(0)"])
            .planNode())
    want = pyoracle.run_plan(plan, [rv])
    rows = want.rows()
    assert len(rows) == int((counts[:6] > 0).sum()) and int(counts[6:].sum()) == 0
    for row in rows:
        g = tpch.RETURNFLAG_DICT.index(row[0]) * 2 + tpch.LINESTATUS_DICT.index(row[1])
        assert int(counts[g]) == row[7], (row, counts)
        for p in range(5):
            assert _close(sums[g, p], row[2 + p], n), (row, p, sums[g])


@pytest.mark.parametrize("n,stages", [(6000, 0), (5 * 1024 + 500, 2)])
def test_q14_probe_side(host, n, stages):
    """Array-mode probe inside the scan: one byte per key slot (0 no build row, 1 match, 2 match with p_type LIKE
    'PROMO%'), rows outside the key range or the date range contribute nothing."""
    grid, nparts = 3, 300
    h = _lineitem(n, seed=14, nparts=nparts)
    part = {k: v.numpy() for k, v in tpch.gen_part(nparts, seed=5).items()}
    # build side: drop a fifth of the parts so that some probes miss, and shrink the table range below the probe keys' range
    keep = np.ones(nparts, dtype=bool)
    keep[::5] = False
    pk, ptype = part["p_partkey"][keep], part["p_type"][keep]
    jmin, jrange = int(pk.min()), int(pk.max() - pk.min() + 1) - 7
    flags = np.zeros(jrange + 1, dtype=np.uint8)
    promo = np.array([tpch.PTYPE_DICT[t].startswith("PROMO") for t in ptype])
    inside = pk - jmin < jrange
    flags[(pk - jmin)[inside]] = np.where(promo[inside], 2, 1)
    a = FusedArgs()
    for c, name in enumerate(["l_shipdate", "l_partkey", "l_extendedprice", "l_discount"]):
        a.cols[c] = _ptr(h[name])
    a.pi[0], a.pi[1] = tpch.Q14_SHIP_LO, tpch.Q14_SHIP_HI
    a.pf[0], a.pf[1], a.pf[2] = 1.0, 1.0, 0.0
    a.rows, a.nkeys, a.ngroups = n, 0, 1
    a.join_slot_flags, a.join_min, a.join_range = _ptr(flags), jmin, jrange
    sums, counts = _run(host, 14, a, grid, 1, 2, stages=stages)
    li_names = ["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"]
    li = row_vector(li_names, [_vec(h, c) for c in li_names])
    pt = row_vector(["p_partkey", "p_type"], [flat_vector(BIGINT, pk[inside]), dictionary_vector(VARCHAR, ptype[inside], tpch.PTYPE_DICT)])
    build_side = PlanBuilder().values(pt.names, pt.types, source=1)
    plan = (PlanBuilder().values(li.names, li.types, source=0).filter("l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE")
            .project(["l_extendedprice * (1.0 - l_discount) as part_revenue", "l_shipdate", "l_partkey"])
            .hashJoin(["l_partkey"], ["p_partkey"], build_side, "", ["part_revenue", "p_type"])
            .project(["(CASE WHEN (p_type LIKE 'PROMO%') THEN part_revenue ELSE 0.0 END) as filter_revenue", "part_revenue"])
            .singleAggregation([], ["sum(part_revenue)", "sum(filter_revenue)", "count(0)"]).planNode())
    want = pyoracle.run_plan(plan, [li, pt])
    ((want_rev, want_promo, want_count),) = want.rows()
    assert int(counts[0]) == want_count and 0 < counts[0] < n
    assert _close(sums[0, 0], want_rev, n) and _close(sums[0, 1], want_promo, n)
    assert 0 < sums[0, 1] < sums[0, 0]


@pytest.mark.parametrize("n,stages,smem_acc", [(2500, 0, 0), (4 * 1024 + 9, 2, 0), (4 * 1024 + 9, 3, 1)])
def test_bigint_key_through_a_lookup_table_and_nan_filter_input(host, n, stages, smem_acc):
    """The generic small shape sum(f1 * (pf1 - f2)) WHERE f0 < pf0 GROUP BY a BIGINT key whose values reach their
    group ids through key_lut[value - key_min] (value-id normalisation of a sparse range, VectorHasher.cpp:560-640);
    NaN in the filter input compares as the largest value (NaN < x is false, velox/type/FloatingPointUtil.h)."""
    grid = 2
    rng = np.random.default_rng(2)
    f0, f1, f2 = rng.random(n), rng.random(n) * 100, rng.random(n)
    f0[::17] = np.nan
    values = np.array([-40, -37, -10, 5], dtype=np.int64)  # four groups scattered over a range of 46
    key = values[rng.integers(0, 4, n)]
    lut = np.full(46, -1, dtype=np.int32)
    lut[values + 40] = [2, 0, 3, 1]
    a = FusedArgs()
    a.cols[0], a.cols[1], a.cols[2] = _ptr(f0), _ptr(f1), _ptr(f2)
    a.pf[0], a.pf[1] = 0.6, 1.0
    a.rows, a.nkeys, a.ngroups = n, 1, 4
    a.key[0], a.key_is64[0], a.key_mult[0], a.key_min[0], a.key_lut[0] = _ptr(key), 1, 1, -40, _ptr(lut)
    sums, counts = _run(host, 2, a, grid, 4, 1, stages=stages, smem_acc=smem_acc)
    keep = f0 < 0.6  # numpy: NaN < x is False as well
    for v, g in zip(values, [2, 0, 3, 1]):
        m = keep & (key == v)
        assert int(counts[g]) == int(m.sum()) and m.sum() > 0
        assert _close(sums[g, 0], float((f1[m] * (1.0 - f2[m])).sum()), n)


def test_scan_compact_in_front_of_the_exchange(host):
    """Multi-GPU Q14: each GPU filters its lineitem shard and ships only (l_partkey, revenue) of the rows in
    the date range. Inside a tile the output follows the input order, tiles land in reservation order: the output is
    the expected rows as a multiset, dense, and nothing is written past `count`. A too small output sets error 100."""
    n, grid, stages = 2 * 3 * 1024 + 400, 2, 2
    h = _lineitem(n, seed=41)
    a = FusedArgs()
    for c, name in enumerate(["l_shipdate", "l_partkey", "l_extendedprice", "l_discount"]):
        a.cols[c] = _ptr(h[name])
    a.pi[0], a.pi[1] = tpch.Q14_SHIP_LO - 200, tpch.Q14_SHIP_HI + 200  # a wider window: a few hundred rows per tile
    a.pf[0] = 1.0
    a.rows = n
    keep = (h["l_shipdate"] >= tpch.Q14_SHIP_LO - 200) & (h["l_shipdate"] <= tpch.Q14_SHIP_HI + 200)
    want = sorted(zip(h["l_partkey"][keep].tolist(), (h["l_extendedprice"][keep] * (1.0 - h["l_discount"][keep])).tolist()))
    assert len(want) > 500
    cap = len(want) + 64
    out_key, out_rev = np.full(cap, -7, dtype=np.int64), np.full(cap, -7.0)
    count, err = np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.int32)
    host.h_compact(C.byref(a), grid, stages, _arg(out_key), _arg(out_rev), C.c_int64(cap), _arg(count), _arg(err))
    assert int(count[0]) == len(want) and int(err[0]) == 0
    assert sorted(zip(out_key[:len(want)].tolist(), out_rev[:len(want)].tolist())) == want
    assert (out_key[len(want):] == -7).all() and (out_rev[len(want):] == -7.0).all()
    # inside a tile the survivors keep their input order: find where tile 0 landed and compare the run
    first = np.nonzero(keep[:1024])[0]
    rev = h["l_extendedprice"] * (1.0 - h["l_discount"])
    (starts,) = np.nonzero((out_key[:len(want)] == h["l_partkey"][first[0]]) & (out_rev[:len(want)] == rev[first[0]]))
    assert any(np.array_equal(out_key[s0:s0 + len(first)], h["l_partkey"][first]) and np.array_equal(out_rev[s0:s0 + len(first)], rev[first])
               for s0 in starts)
    short = len(want) - 100
    out_key2, out_rev2 = np.full(short, -7, dtype=np.int64), np.full(short, -7.0)
    count[0], err[0] = 0, 0
    host.h_compact(C.byref(a), grid, stages, _arg(out_key2), _arg(out_rev2), C.c_int64(short), _arg(count), _arg(err))
    assert int(err[0]) == 100 and int(count[0]) == len(want)  # the count still says how much room was needed


def test_late_materialisation_filter_bitmap_then_gather(host):
    """Selective filters (FilterProject.cpp:200-259: the projections only see the surviving rows): the filter's one
    4-byte column goes through the ring in 4096-row tiles and becomes a selection bitmap; every other column is touched
    only at the surviving row numbers by the gather kernel, which probes, projects and aggregates. With
    tile_stride = 2 the same kernel samples every second tile and only counts (the planner's selectivity estimate)."""
    grid, stages, nparts = 2, 2, 300
    tile = 4096
    n = 2 * 3 * tile + 1000 + 13  # three tiles per block (ring of two: wraps), a tail of whole and partial words
    h = _lineitem(n, seed=77, nparts=nparts)
    a = FusedArgs()
    for c, name in enumerate(["l_shipdate", "l_partkey", "l_extendedprice", "l_discount"]):
        a.cols[c] = _ptr(h[name])
    a.pi[0], a.pi[1] = tpch.Q14_SHIP_LO, tpch.Q14_SHIP_HI
    a.pf[0], a.pf[1], a.pf[2] = 1.0, 1.0, 0.0
    a.rows = n
    keep = (h["l_shipdate"] >= tpch.Q14_SHIP_LO) & (h["l_shipdate"] <= tpch.Q14_SHIP_HI)
    bits = np.full((n + 31) // 32 + 4, 0xDEADBEEF, dtype=np.uint32)
    counters = np.zeros(2, dtype=np.uint64)
    assert host.h_filter_bits(C.byref(a), grid, stages, 1, _arg(bits), _arg(counters)) == tile
    nwords = (n + 31) // 32
    want_bits = np.zeros(nwords * 32, dtype=bool)
    want_bits[:n] = keep
    want_words = np.packbits(want_bits.reshape(-1, 32), axis=1, bitorder="little").view(np.uint32).ravel()
    assert np.array_equal(bits[:nwords], want_words) and (bits[nwords:] == 0xDEADBEEF).all()
    assert int(counters[0]) == int(keep.sum()) > 0 and int(counters[1]) == n
    # sampling pass: tiles 0, 2, 4 of the six whole tiles, no bitmap, no tail
    counters[:] = 0
    host.h_filter_bits(C.byref(a), grid, stages, 2, None, _arg(counters))
    sampled = np.concatenate([keep[t * tile:(t + 1) * tile] for t in (0, 2, 4)])
    assert int(counters[0]) == int(sampled.sum()) and int(counters[1]) == 3 * tile
    # gather over the surviving rows; build side as in test_q14_probe_side
    part = {k: v.numpy() for k, v in tpch.gen_part(nparts, seed=5).items()}
    live = np.ones(nparts, dtype=bool)
    live[::4] = False
    pk, ptype = part["p_partkey"][live], part["p_type"][live]
    jmin, jrange = int(pk.min()), int(pk.max() - pk.min() + 1)
    flags = np.zeros(jrange, dtype=np.uint8)
    flags[pk - jmin] = np.where([tpch.PTYPE_DICT[t].startswith("PROMO") for t in ptype], 2, 1)
    a.join_slot_flags, a.join_min, a.join_range = _ptr(flags), jmin, jrange
    sel = np.nonzero(keep)[0].astype(np.int32)
    nsel = np.array([len(sel)], dtype=np.int64)
    sums, counts = np.zeros(2), np.zeros(1, dtype=np.int64)
    host.h_gather(C.byref(a), 3, _arg(sel), _arg(nsel), _arg(sums), _arg(counts))
    li_names = ["l_partkey", "l_extendedprice", "l_discount", "l_shipdate"]
    li = row_vector(li_names, [_vec(h, c) for c in li_names])
    pt = row_vector(["p_partkey", "p_type"], [flat_vector(BIGINT, pk), dictionary_vector(VARCHAR, ptype, tpch.PTYPE_DICT)])
    build_side = PlanBuilder().values(pt.names, pt.types, source=1)
    plan = (PlanBuilder().values(li.names, li.types, source=0).filter("l_shipdate between '1995-09-01'::DATE and '1995-09-30'::DATE")
            .project(["l_extendedprice * (1.0 - l_discount) as part_revenue", "l_shipdate", "l_partkey"])
            .hashJoin(["l_partkey"], ["p_partkey"], build_side, "", ["part_revenue", "p_type"])
            .project(["(CASE WHEN (p_type LIKE 'PROMO%') THEN part_revenue ELSE 0.0 END) as filter_revenue", "part_revenue"])
            .singleAggregation([], ["sum(part_revenue)", "sum(filter_revenue)", "count(0)"]).planNode())
    ((want_rev, want_promo, want_count),) = pyoracle.run_plan(plan, [li, pt]).rows()
    assert int(counts[0]) == want_count and 0 < want_count < len(sel)
    assert _close(sums[0], want_rev, n) and _close(sums[1], want_promo, n)


# ---- pipelines outside the ahead-of-time list: signature -> expression-template type (fused_jit.cu) -> kernel ----------------
JIT_SIGS = [
    "F:and(between(i0,pi0,pi1),between(f1,pf0,pf1),lt(f2,pf2),gt(f3,pf3));P:multiply(f3,f1)",
    "F:lt(f0,pf0);P:divide(f1,plus(pf1,f0))|minus(f0,f1)",
    "F:and(neq(l0,pl0),lte(f1,pf0));P:switch(gte(f1,pf1),f2,pf2)",
    "F:eq(i0,pi0);P:f2|switch(joinflag,multiply(f2,pf0),pf1);J:l1",
    "F:true;P:plus(f0,f1)",
]


def _parse(s, p=0):
    b = p
    while p < len(s) and (s[p].isalnum() or s[p] == "_"):
        p += 1
    name, args = s[b:p], []
    if p < len(s) and s[p] == "(":
        p += 1
        while True:
            arg, p = _parse(s, p)
            args.append(arg)
            if s[p] == ",":
                p += 1
                continue
            assert s[p] == ")"
            p += 1
            break
    return (name, args), p


def _lt(a, b):  # NaN sorts above every number (velox/type/FloatingPointUtil.h)
    return (a < b) | (~np.isnan(a) & np.isnan(b))


def _eq(a, b):
    return (a == b) | (np.isnan(a) & np.isnan(b))


def _ev(node, env):
    """An independent evaluator of the signature text over numpy columns."""
    name, args = node
    if not args:
        if name == "true":
            return np.ones(env["n"], dtype=bool)
        if name == "joinflag":
            return env["joinflag"]
        if name[0] == "p":
            return env["consts"][name[1]][int(name[2:])]
        return env["cols"][int(name[1:])]
    v = [_ev(x, env) for x in args]
    with np.errstate(all="ignore"):
        return {"and": lambda: np.logical_and.reduce(v), "between": lambda: ~_lt(v[0], v[1]) & ~_lt(v[2], v[0]),
                "lt": lambda: _lt(v[0], v[1]), "gt": lambda: _lt(v[1], v[0]), "lte": lambda: ~_lt(v[1], v[0]), "gte": lambda: ~_lt(v[0], v[1]),
                "eq": lambda: _eq(v[0], v[1]), "neq": lambda: ~_eq(v[0], v[1]), "plus": lambda: v[0] + v[1], "minus": lambda: v[0] - v[1],
                "multiply": lambda: v[0] * v[1], "divide": lambda: v[0] / v[1], "switch": lambda: np.where(v[0], v[1], v[2])}[name]()


def _top_level_split(text, sep):
    out, depth, b = [], 0, 0
    for i, ch in enumerate(text):
        depth += (ch == "(") - (ch == ")")
        if ch == sep and depth == 0:
            out.append(text[b:i])
            b = i + 1
    return out + [text[b:]]


@pytest.fixture(scope="module")
def jit_host(tmp_path_factory):
    """The library's signature parser (vb2k_pipeline_jit_compiles: also proves each instantiation compiles for sm_100a
    under NVRTC) hands back the C++ type it generated; that text becomes `using J<i> = ...` in the host build."""
    from velox_b200._lib import lib
    L = lib()
    L.vb2k_pipeline_jit_compiles.argtypes = [C.c_char_p, C.c_int32, C.c_int32, C.c_int32, C.c_char_p, C.c_int32]
    types = []
    for sig in JIT_SIGS:
        buf = C.create_string_buffer(4000)
        assert L.vb2k_pipeline_jit_compiles(sig.encode(), 1, 1, 0, buf, 4000) == 1, (sig, buf.value.decode()[:1500])
        types.append(buf.value.decode())
    extra_types = "".join(f"using J{i} = {t};\n" for i, t in enumerate(types))
    cases = "".join(f"    case {i}: return run<J{i}, 1, int32_t>(a, grid, stages, sums, counts);\n" for i in range(len(types)))
    sigs = "".join(f"    case {i}: s = J{i}::sig(); break;\n" for i in range(len(types)))
    drivers = """
int h_jit(int which, const vb2_fused_args* in, int grid, int stages, double* sums, int64_t* counts) {
  KernelArgs a{};
  for (int c = 0; c < kMaxCols; ++c) a.cols[c] = in->cols[c];
  for (int k = 0; k < VB2_FUSED_MAX_PARAMS; ++k) { a.consts.pf[k] = in->pf[k]; a.consts.pl[k] = in->pl[k]; a.consts.pi[k] = in->pi[k]; }
  a.rows = in->rows;
  a.ngroups = 1;
  a.join_slot_flags = reinterpret_cast<const uint8_t*>(in->join_slot_flags);
  a.join_min = in->join_min;
  a.join_range = in->join_range;
  a.release_guard = kReleaseGuard;
  switch (which) {
%s  }
  return 1;
}
const char* h_jit_sig(int which) {
  static std::string s;
  switch (which) {
%s  }
  return s.c_str();
}
""" % (cases, sigs)
    H = _build(tmp_path_factory.mktemp("fused_jit_on_host"), extra_types, drivers)
    H.h_jit_sig.restype = C.c_char_p
    return H


@pytest.mark.parametrize("which", range(len(JIT_SIGS)))
@pytest.mark.parametrize("stages,with_nan", [(0, False), (0, True), (2, False)])
def test_jit_pipeline_types_round_trip_and_evaluate(jit_host, which, stages, with_nan):
    """Plan shapes without an ahead-of-time specialisation are instantiated at run time from the signature the planner
    prints (host/fused_match.cpp -> fused_jit.cu parse_signature). Here: the generated type prints the signature it was
    parsed from (parser and printers are inverse), and running it -- direct-load and TMA-staged -- gives what an
    independent numpy evaluation of the signature text gives, NaN inputs included."""
    sig = JIT_SIGS[which]
    assert jit_host.h_jit_sig(which).decode() == sig
    n, grid = 3 * 1024 + 333, 2
    rng = np.random.default_rng(100 + which)
    body, _, join = sig[2:].partition(";J:l")
    ftext, _, ptext = body.partition(";P:")
    names = set(re.findall(r"\b([fil])(\d+)\b", sig.replace(";J:l", ";J:,l")))
    cols, keep_alive = {}, []
    a = FusedArgs()
    for kind, idx in sorted(names):
        c = int(idx)
        if kind == "f":
            col = rng.random(n) * 10
            if with_nan:
                col[rng.random(n) < 0.02] = np.nan
        elif kind == "i":
            col = rng.integers(0, 4, n).astype(np.int32)
        else:
            col = rng.integers(0, 50, n).astype(np.int64)
        assert c not in cols
        cols[c] = col
        a.cols[c] = _ptr(col)
    consts = {"f": [3.0, 6.0, 5.0, 2.0] + [0.0] * 8, "i": [1, 2] + [0] * 10, "l": [7] + [0] * 11}
    for k in range(12):
        a.pf[k], a.pi[k], a.pl[k] = consts["f"][k], consts["i"][k], consts["l"][k]
    a.rows = n
    env = {"n": n, "cols": cols, "consts": consts, "joinflag": None}
    keep = _ev(_parse(ftext)[0], env)
    if join:
        jmin, jrange = 5, 40  # probe keys 0..49: some below, some above the table
        flags = rng.integers(0, 3, jrange).astype(np.uint8)
        keep_alive.append(flags)
        a.join_slot_flags, a.join_min, a.join_range = _ptr(flags), jmin, jrange
        slot = cols[int(join)] - jmin
        inside = (slot >= 0) & (slot < jrange)
        hit = np.where(inside, flags[np.clip(slot, 0, jrange - 1)], 0)
        keep = keep & (hit != 0)
        env["joinflag"] = hit == 2
    projs = _top_level_split(ptext, "|")
    sums = np.zeros(len(projs))
    counts = np.zeros(1, dtype=np.int64)
    assert jit_host.h_jit(which, C.byref(a), grid, stages, _arg(sums), _arg(counts)) == 0
    assert int(counts[0]) == int(keep.sum()) and 0 < keep.sum() < n + (ftext == "true")
    for p, text in enumerate(projs):
        want = np.broadcast_to(_ev(_parse(text)[0], env), (n,))[keep]
        if np.isnan(want).any():
            assert with_nan and np.isnan(sums[p]), (text, sums[p])
        else:
            assert _close(sums[p], float(want.sum()), n), (text, sums[p], want.sum())
