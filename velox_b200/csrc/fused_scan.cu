// Registry, launch logic and the ahead-of-time specialised pipelines of the fused scan path.
#include "fused_scan.cuh"

#include <cstring>
#include <deque>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace vb2 {
namespace fx {

// One warp per accumulator: lanes stride over the per-block partials, then a fixed butterfly.
__global__ void fused_finalize_kernel(const double* __restrict__ partials, int nblocks, int kvals, int np, int maxg,
                                      int ngroups, double* __restrict__ sums, int64_t* __restrict__ counts) {
  const int t = blockIdx.x;
  const int lane = threadIdx.x;
  const int g = t / (np + 1), p = t % (np + 1);
  if (t >= kvals || g >= ngroups || g >= maxg) return;
  if (p == np) {
    int64_t c = 0;
    for (int b = lane; b < nblocks; b += 32) c += __double_as_longlong(partials[static_cast<int64_t>(b) * kvals + t]);
    c = warp_sum(c);
    if (lane == 0) counts[g] += c;
  } else {
    double s = 0.0;
    for (int b = lane; b < nblocks; b += 32) s = __dadd_rn(s, partials[static_cast<int64_t>(b) * kvals + t]);
    s = warp_sum(s);
    if (lane == 0) sums[g * np + p] = __dadd_rn(sums[g * np + p], s);
  }
}

// join_slot_flags[slot] = 0 (no build row) | 1 (match) | 2 (match and build-side predicate true)
__global__ void join_slot_flags_kernel(const int32_t* __restrict__ head, const int32_t* __restrict__ codes,
                                       const uint8_t* __restrict__ flag, int64_t range, uint8_t* __restrict__ out) {
  for (int64_t s = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; s < range; s += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const int32_t h = head[s];
    uint8_t v = 0;
    if (h != 0) {
      const int32_t code = codes ? codes[h - 1] : (h - 1);
      v = (flag && flag[code]) ? 2 : 1;
    }
    out[s] = v;
  }
}

static std::deque<Entry>& registry() {  // deque: entries keep their addresses when pipelines are added at run time
  static std::deque<Entry> r;
  return r;
}
static std::mutex& registry_mutex() {
  static std::mutex m;
  return m;
}

int register_pipeline(const Entry& e) {
  std::lock_guard<std::mutex> lock(registry_mutex());
  registry().push_back(e);
  return static_cast<int>(registry().size()) - 1;
}

constexpr int kMaxBlocksPerSM = 8;
constexpr int kFusedMaxGroups = 64;       // upper bound; the real limit is shared-memory capacity
constexpr size_t kSmemLimit = 220 * 1024;  // of the 227 KB a block may use

static int run_finalize(const KernelArgs& a, void* ws, int64_t grid, int kvals, int np, int maxg, double* sums, int64_t* counts,
                        cudaStream_t st) {
  fused_finalize_kernel<<<vb2::counted(kvals), 32, 0, st>>>(reinterpret_cast<const double*>(ws), static_cast<int>(grid), kvals, np, maxg, a.ngroups, sums, counts);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

static int popc32(uint32_t m) { int n = 0; for (; m; m &= m - 1) ++n; return n; }
// TileLayout<P, kKeyBytes>::stage_bytes(nkeys) from the run-time column masks
static int stage_bytes_of(uint32_t fmask, uint32_t imask, uint32_t lmask, int nkeys, int key_bytes, int kTileRows = ::vb2::fx::kTileRows) {
  const int end = 8 * kTileRows * (popc32(fmask) + popc32(lmask)) + (key_bytes == 8 ? nkeys * 8 * kTileRows : 0) + 4 * kTileRows * popc32(imask) +
                  (key_bytes == 4 ? nkeys * 4 * kTileRows : 0);
  return (end + 127) / 128 * 128;
}

// Occupancy and the dynamic shared-memory opt-in are per kernel and do not change: remember them.
struct KernelSetup {
  size_t smem_configured = 0;
  int blocks_per_sm = 0;
  size_t blocks_for_smem = static_cast<size_t>(-1);
};
static std::mutex g_setup_mu;
static std::unordered_map<const void*, KernelSetup> g_setup;

static int blocks_per_sm_of(const void* kernel, int threads, size_t smem, int* out) {
  std::lock_guard<std::mutex> lock(g_setup_mu);
  KernelSetup& ks = g_setup[kernel];
  if (ks.smem_configured < smem) {
    VB2_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    ks.smem_configured = smem;
  }
  if (ks.blocks_for_smem != smem) {
    int n = 0;
    VB2_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem));
    ks.blocks_per_sm = n < 1 ? 1 : (n > kMaxBlocksPerSM ? kMaxBlocksPerSM : n);
    ks.blocks_for_smem = smem;
  }
  *out = ks.blocks_per_sm;
  return VB2_OK;
}

static int launch_kernel(const void* fn, unsigned grid, unsigned block, size_t smem, void** params, cudaStream_t st) {
  note_launch();
  VB2_CUDA_OK(cudaLaunchKernel(fn, dim3(grid), dim3(block), params, smem, st));
  return VB2_OK;
}

static bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

// Picks the accumulator variant and the data path for one batch and launches it.
static int launch_aggregate(const Entry& e, KernelArgs a, double* sums, int64_t* counts, void* ws, size_t ws_bytes, cudaStream_t st) {
  const PipelineDesc& d = e.desc;
  bool bulk = true;  // cp.async.bulk needs 16-byte aligned sources (inputs below one tile run as the kernel's tail)
  for (int c = 0; c < kMaxCols; ++c)
    if (((d.fmask | d.imask | d.lmask) >> c) & 1u) bulk = bulk && aligned16(a.cols[c]);
  for (int k = 0; k < a.nkeys; ++k) bulk = bulk && aligned16(a.key[k]);
  const int g = a.nkeys == 0 ? 1 : a.ngroups;
  if (g > kFusedMaxGroups) return fail_msg(VB2_ERR_UNSUPPORTED, "fused aggregation: group-id space too large");
  bool key64 = false, key32 = false;
  for (int k = 0; k < a.nkeys; ++k) (a.key_is64[k] ? key64 : key32) = true;
  if (key64 && key32) return fail_msg(VB2_ERR_UNSUPPORTED, "fused aggregation: group keys must have one width");
  int maxg;  // 0 = shared-memory accumulators
  if (g <= 1) { maxg = 1; key64 = false; }
  else if (g <= 4) maxg = 4;
  else if (bulk) maxg = 0;
  else if (g <= 8) maxg = 8;
  else return fail_msg(VB2_ERR_UNSUPPORTED, "fused aggregation: unaligned input with more than 8 groups");
  const int np = d.nproj;
  if (bulk) {
    const void* fn = e.kernels(e.self, KernelKind::kTma, maxg, key64);
    if (!fn) return fail_msg(VB2_ERR_UNSUPPORTED, "fused aggregation: kernel variant unavailable");
    const bool smem_acc = maxg == 0;
    const int stage_bytes = stage_bytes_of(d.fmask, d.imask, d.lmask, (smem_acc || maxg > 1) ? a.nkeys : 0, key64 ? 8 : 4);
    const size_t acc_bytes = smem_acc ? static_cast<size_t>(a.ngroups + 1) * kConsumerThreads * (np * 8 + 4) : 0;
    const int budget = smem_acc ? static_cast<int>(kSmemLimit - acc_bytes) : 100 * 1024;
    int stages = budget / stage_bytes;
    if (stages < 2) return fail_msg(VB2_ERR_UNSUPPORTED, "fused aggregation: too many groups for shared-memory accumulators");
    stages = stages > kMaxStages ? kMaxStages : stages;
    const size_t smem = static_cast<size_t>(stages) * stage_bytes + acc_bytes;
    int bps = 1;
    if (int rc = blocks_per_sm_of(fn, kTmaThreads, smem, &bps)) return rc;
    const int64_t ntiles = a.rows / kTileRows;
    int64_t grid = static_cast<int64_t>(device_sm_count()) * bps;
    if (ntiles < grid) grid = ntiles < 1 ? 1 : ntiles;
    const int groups = smem_acc ? a.ngroups : maxg;
    const int kvals = groups * (np + 1);
    if (ws_bytes < static_cast<size_t>(grid) * kvals * sizeof(double)) return fail_msg(VB2_ERR_INVALID, "fused workspace too small");
    double* partials = reinterpret_cast<double*>(ws);
    void* params[] = {&a, &stages, &partials};
    if (int rc = launch_kernel(fn, static_cast<unsigned>(grid), kTmaThreads, smem, params, st)) return rc;
    return run_finalize(a, ws, grid, kvals, np, groups, sums, counts, st);
  }
  const void* fn = e.kernels(e.self, KernelKind::kDirect, maxg, key64);
  if (!fn) return fail_msg(VB2_ERR_UNSUPPORTED, "fused aggregation: kernel variant unavailable");
  int bps = 1;
  if (int rc = blocks_per_sm_of(fn, kThreads, 0, &bps)) return rc;
  int64_t want = (a.rows + kThreads - 1) / kThreads;
  int64_t grid = static_cast<int64_t>(device_sm_count()) * bps;
  if (want < grid) grid = want < 1 ? 1 : want;
  const int kvals = maxg * (np + 1);
  if (ws_bytes < static_cast<size_t>(grid) * kvals * sizeof(double)) return fail_msg(VB2_ERR_INVALID, "fused workspace too small");
  double* partials = reinterpret_cast<double*>(ws);
  void* params[] = {&a, &partials};
  if (int rc = launch_kernel(fn, static_cast<unsigned>(grid), kThreads, 0, params, st)) return rc;
  return run_finalize(a, ws, grid, kvals, np, maxg, sums, counts, st);
}

// Selection bitmap of the pipeline's filter over a.rows rows (tile_stride == 1), or a strided sample
// that only counts (bits == nullptr). counters: device {rows kept, rows evaluated}, accumulated.
static int launch_filter_bits(const Entry& e, KernelArgs a, int tile_stride, uint32_t* bits, unsigned long long* counters, cudaStream_t st) {
  const PipelineDesc& d = e.desc;
  if (!d.has_filter) return fail_msg(VB2_ERR_UNSUPPORTED, "fused pipeline without a filter");
  for (int c = 0; c < kMaxCols; ++c)
    if ((((d.ffmask | d.fimask | d.flmask) >> c) & 1u) && !aligned16(a.cols[c])) return fail_msg(VB2_ERR_UNSUPPORTED, "fused filter needs 16-byte aligned columns");
  const void* fn = e.kernels(e.self, KernelKind::kFilterBits, 0, false);
  if (!fn) return fail_msg(VB2_ERR_UNSUPPORTED, "fused filter: kernel unavailable");
  const int tile_rows = filter_tile_rows_for(d.ffmask, d.fimask, d.flmask);
  const int stage_bytes = stage_bytes_of(d.ffmask, d.fimask, d.flmask, 0, 4, tile_rows);
  int stages = (64 * 1024) / stage_bytes;
  stages = stages > kMaxStages ? kMaxStages : (stages < 2 ? 2 : stages);
  const size_t smem = static_cast<size_t>(stages) * stage_bytes;
  int bps = 1;
  if (int rc = blocks_per_sm_of(fn, kTmaThreads, smem, &bps)) return rc;
  if (tile_stride < 1) tile_stride = 1;
  const int64_t ntiles = (a.rows / tile_rows + tile_stride - 1) / tile_stride;
  int64_t grid = static_cast<int64_t>(device_sm_count()) * bps;
  if (ntiles < grid) grid = ntiles < 1 ? 1 : ntiles;
  void* params[] = {&a, &stages, &tile_stride, &bits, &counters};
  return launch_kernel(fn, static_cast<unsigned>(grid), kTmaThreads, smem, params, st);
}

// Probe + project + aggregate over the selected rows sel[0 .. *nsel_dev); nsel_hint sizes the grid.
static int launch_gather(const Entry& e, KernelArgs a, const int32_t* sel, const int64_t* nsel_dev, int64_t nsel_hint, double* sums, int64_t* counts,
                         void* ws, size_t ws_bytes, cudaStream_t st) {
  const PipelineDesc& d = e.desc;
  const int g = a.nkeys == 0 ? 1 : a.ngroups;
  bool key64 = false, key32 = false;
  for (int k = 0; k < a.nkeys; ++k) (a.key_is64[k] ? key64 : key32) = true;
  if (g > 4 || (key64 && key32)) return fail_msg(VB2_ERR_UNSUPPORTED, "fused gather: at most 4 groups, one key width");
  const int maxg = g <= 1 ? 1 : 4;
  if (maxg == 1) key64 = false;
  const void* fn = e.kernels(e.self, KernelKind::kGather, maxg, key64);
  if (!fn) return fail_msg(VB2_ERR_UNSUPPORTED, "fused gather: kernel variant unavailable");
  int bps = 1;
  if (int rc = blocks_per_sm_of(fn, kThreads, 0, &bps)) return rc;
  int64_t want = (nsel_hint + kThreads * 4 - 1) / (kThreads * 4);
  int64_t grid = static_cast<int64_t>(device_sm_count()) * bps;
  if (want < grid) grid = want < 1 ? 1 : want;
  const int kvals = maxg * (d.nproj + 1);
  if (ws_bytes < static_cast<size_t>(grid) * kvals * sizeof(double)) return fail_msg(VB2_ERR_INVALID, "fused workspace too small");
  double* partials = reinterpret_cast<double*>(ws);
  void* params[] = {&a, &sel, &nsel_dev, &partials};
  if (int rc = launch_kernel(fn, static_cast<unsigned>(grid), kThreads, 0, params, st)) return rc;
  return run_finalize(a, ws, grid, kvals, d.nproj, maxg, sums, counts, st);
}

// ---- ahead-of-time instantiations behind the same interface ------------------------------------------
template <class P>
static const void* aot_kernels(void*, KernelKind kind, int maxg, bool key64) {
  switch (kind) {
    case KernelKind::kTma:
      if (maxg == 1) return reinterpret_cast<const void*>(&fused_scan_agg_tma_kernel<P, 1, int32_t>);
      if (maxg == 4) return key64 ? reinterpret_cast<const void*>(&fused_scan_agg_tma_kernel<P, 4, int64_t>) : reinterpret_cast<const void*>(&fused_scan_agg_tma_kernel<P, 4, int32_t>);
      if (maxg == 0) return key64 ? reinterpret_cast<const void*>(&fused_scan_agg_tma_kernel<P, 0, int64_t>) : reinterpret_cast<const void*>(&fused_scan_agg_tma_kernel<P, 0, int32_t>);
      return nullptr;
    case KernelKind::kDirect:
      if (maxg == 1) return reinterpret_cast<const void*>(&fused_scan_agg_kernel<P, 1, 2, false, int32_t>);
      if (maxg == 4) return key64 ? reinterpret_cast<const void*>(&fused_scan_agg_kernel<P, 4, 2, false, int64_t>) : reinterpret_cast<const void*>(&fused_scan_agg_kernel<P, 4, 2, false, int32_t>);
      if (maxg == 8) return key64 ? reinterpret_cast<const void*>(&fused_scan_agg_kernel<P, 8, 2, false, int64_t>) : reinterpret_cast<const void*>(&fused_scan_agg_kernel<P, 8, 2, false, int32_t>);
      return nullptr;
    case KernelKind::kFilterBits:
      if constexpr (!is_same_v<typename P::F, True>)
        return reinterpret_cast<const void*>(&fused_filter_bits_tma_kernel<typename P::FilterView, filter_tile_rows_for(P::F::fmask, P::F::imask, P::F::lmask)>);
      return nullptr;
    case KernelKind::kGather:
      if constexpr (!is_same_v<typename P::F, True>) {
        if (maxg == 1) return reinterpret_cast<const void*>(&fused_gather_agg_kernel<P, 1, int32_t>);
        if (maxg == 4) return key64 ? reinterpret_cast<const void*>(&fused_gather_agg_kernel<P, 4, int64_t>) : reinterpret_cast<const void*>(&fused_gather_agg_kernel<P, 4, int32_t>);
      }
      return nullptr;
  }
  return nullptr;
}

template <class P>
static PipelineDesc describe_pipeline() {
  PipelineDesc d;
  d.nproj = P::kNP;
  d.join = P::kJoin;
  d.has_filter = !is_same_v<typename P::F, True>;
  d.fmask = P::fmask; d.imask = P::imask; d.lmask = P::lmask;
  d.ffmask = P::F::fmask; d.fimask = P::F::imask; d.flmask = P::F::lmask;
  return d;
}

template <class P>
static int add_pipeline() {
  Entry e;
  e.signature = P::sig();
  e.nproj = P::kNP;
  e.join = P::kJoin;
  e.desc = describe_pipeline<P>();
  e.kernels = &aot_kernels<P>;
  return register_pipeline(e);
}

// A filter on its own ("F:<filter>;P:"): what B200FilterProject runs for a filter that is not
// absorbed into an aggregation.
template <class Filter>
struct FilterOnly {
  using F = Filter;
  struct FilterView {
    static constexpr uint32_t fmask = Filter::fmask, imask = Filter::imask, lmask = Filter::lmask;
    using F = Filter;
  };
  static std::string sig() { return "F:" + Filter::sig() + ";P:"; }
};
template <class FO>
static const void* aot_filter_only(void*, KernelKind kind, int, bool) {
  return kind == KernelKind::kFilterBits
             ? reinterpret_cast<const void*>(&fused_filter_bits_tma_kernel<typename FO::FilterView, filter_tile_rows_for(FO::F::fmask, FO::F::imask, FO::F::lmask)>)
             : nullptr;
}
template <class FO>
static int add_filter_only() {
  Entry e;
  e.signature = FO::sig();
  e.desc.has_filter = true;
  e.desc.ffmask = e.desc.fmask = FO::F::fmask;
  e.desc.fimask = e.desc.imask = FO::F::imask;
  e.desc.flmask = e.desc.lmask = FO::F::lmask;
  e.kernels = &aot_filter_only<FO>;
  return register_pipeline(e);
}

template <class P>
static int launch_compact(const KernelArgs& a, const CompactArgs& o, cudaStream_t st) {
  for (int c = 0; c < kMaxCols; ++c)
    if ((((P::fmask | P::imask | P::lmask) >> c) & 1u) && !aligned16(a.cols[c]))
      return fail_msg(VB2_ERR_UNSUPPORTED, "fused scan-compact needs 16-byte aligned input columns");
  auto kernel = fused_scan_compact_tma_kernel<P>;
  const int stage_bytes = TileLayout<P, 4>::stage_bytes(0);
  int stages = (100 * 1024) / stage_bytes;
  stages = stages > kMaxStages ? kMaxStages : (stages < 2 ? 2 : stages);
  const size_t smem = static_cast<size_t>(stages) * stage_bytes;
  static size_t configured = 0;
  if (configured < smem) {
    VB2_CUDA_OK(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    configured = smem;
  }
  int blocks_per_sm = 0;
  VB2_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&blocks_per_sm, kernel, kTmaThreads, smem));
  if (blocks_per_sm < 1) blocks_per_sm = 1;
  if (blocks_per_sm > kMaxBlocksPerSM) blocks_per_sm = kMaxBlocksPerSM;
  const int64_t ntiles = a.rows / kTileRows;
  int64_t grid = static_cast<int64_t>(device_sm_count()) * blocks_per_sm;
  if (ntiles < grid) grid = ntiles < 1 ? 1 : ntiles;
  kernel<<<vb2::counted(static_cast<unsigned>(grid)), kTmaThreads, smem, st>>>(a, o, stages);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

template <class P>
static int add_compact_pipeline() {
  Entry e;
  e.signature = P::sig();
  e.nproj = P::kNP;
  e.compact = &launch_compact<P>;
  P::widths(e.widths);
  return register_pipeline(e);
}

// ---------------------------------------------------------------------------------------------
// Specialised pipelines. Column / constant numbering = first use in a depth-first walk of the
// filter, then the join key, then the aggregate-input expressions (see host/fused_match.cpp).
// ---------------------------------------------------------------------------------------------
// TPC-H Q6 (exec/tests/utils/TpchQueryBuilder.cpp:756-788):
//   l_shipdate between d0 and d1 and l_discount between 0.05 and 0.07 and l_quantity < 24.0
//   sum(l_extendedprice * l_discount)
using Q6 = Pipeline<And<Between<ColI<0>, PI<0>, PI<1>>, Between<ColF<1>, PF<0>, PF<1>>, Lt<ColF<2>, PF<2>>>,
                    TypeList<Multiply<ColF<3>, ColF<1>>>>;

// TPC-H Q1 (TpchQueryBuilder.cpp:203-256): l_shipdate < d; aggregate inputs l_quantity,
// l_extendedprice, ep*(1-disc), ep*(1-disc)*(1+tax), l_discount.
using Q1 = Pipeline<Lt<ColI<0>, PI<0>>,
                    TypeList<ColF<1>, ColF<2>, Multiply<ColF<2>, Minus<PF<0>, ColF<3>>>,
                             Multiply<Multiply<ColF<2>, Minus<PF<1>, ColF<3>>>, Plus<PF<2>, ColF<4>>>, ColF<3>>>;

// TPC-H Q14 probe side (TpchQueryBuilder.cpp:1639-1702): l_shipdate between d0 and d1, probe
// l_partkey, sum(ep*(1-disc)), sum(case when p_type like 'PROMO%' then ep*(1-disc) else 0.0 end).
using Q14 = Pipeline<Between<ColI<0>, PI<0>, PI<1>>,
                     TypeList<Multiply<ColF<2>, Minus<PF<0>, ColF<3>>>,
                              Switch<JoinFlag, Multiply<ColF<2>, Minus<PF<1>, ColF<3>>>, PF<2>>>,
                     1>;

// Generic small shapes: sum of one column / product under a single range or comparison filter.
using SumUnderLt = Pipeline<Lt<ColF<0>, PF<0>>, TypeList<Multiply<ColF<1>, Minus<PF<1>, ColF<2>>>>>;
using SumNoFilter = Pipeline<True, TypeList<ColF<0>>>;

// Multi-GPU Q14: each GPU filters its lineitem shard and compacts
// (l_partkey, ep*(1-disc)) for the hash-partitioned exchange ...
using Q14ScanCompact = CompactPipeline<Between<ColI<0>, PI<0>, PI<1>>, TypeList<ColL<1>, Multiply<ColF<2>, Minus<PF<0>, ColF<3>>>>>;
// ... and after the all-to-all probes its part partition: sum(rev), sum(case when promo then rev else 0.0).
using Q14ProbeAfterExchange = Pipeline<True, TypeList<ColF<1>, Switch<JoinFlag, ColF<1>, PF<0>>>, 0>;

static std::once_flag g_once;
static void ensure_registered() {
  std::call_once(g_once, [] {
    add_pipeline<Q6>();
    add_pipeline<Q1>();
    add_pipeline<Q14>();
    add_pipeline<SumUnderLt>();
    add_pipeline<SumNoFilter>();
    add_compact_pipeline<Q14ScanCompact>();
    add_pipeline<Q14ProbeAfterExchange>();
    add_filter_only<FilterOnly<Between<ColI<0>, PI<0>, PI<1>>>>();
    add_filter_only<FilterOnly<Lt<ColI<0>, PI<0>>>>();
  });
}

// fused_jit.cu: NVRTC instantiation of the same templates for signatures without an entry
int jit_pipeline(const std::string& signature);

}  // namespace fx
}  // namespace vb2

using namespace vb2;
using namespace vb2::fx;

extern "C" {

static int find_registered(const char* signature) {
  std::lock_guard<std::mutex> lock(registry_mutex());
  auto& r = registry();
  for (size_t i = 0; i < r.size(); ++i)
    if (r[i].signature == signature) return static_cast<int>(i);
  return -1;
}
int vb2k_fused_find(const char* signature) {
  ensure_registered();
  int id = find_registered(signature);
  if (id >= 0) return id;
  // no ahead-of-time specialisation: instantiate the same templates for this shape with NVRTC
  // (kernels are compiled lazily, per variant, at their first launch)
  static std::mutex jit_mu;
  std::lock_guard<std::mutex> lock(jit_mu);
  id = find_registered(signature);
  if (id >= 0) return id;
  return jit_pipeline(signature);
}
int vb2k_fused_count(void) {
  ensure_registered();
  return static_cast<int>(registry().size());
}
const char* vb2k_fused_signature(int32_t id) {
  ensure_registered();
  std::lock_guard<std::mutex> lock(registry_mutex());
  if (id < 0 || id >= static_cast<int>(registry().size())) return nullptr;
  return registry()[id].signature.c_str();
}
int32_t vb2k_fused_nproj(int32_t id) {
  ensure_registered();
  if (id < 0 || id >= static_cast<int>(registry().size())) return -1;
  return registry()[id].nproj;
}
size_t vb2k_fused_workspace_bytes(int32_t id, int32_t ngroups) {
  ensure_registered();
  if (id < 0 || id >= static_cast<int>(registry().size())) return 0;
  const int g = ngroups <= 1 ? 1 : (ngroups <= 4 ? 4 : (ngroups < 8 ? 8 : ngroups));
  return static_cast<size_t>(device_sm_count()) * kMaxBlocksPerSM * g * (registry()[id].nproj + 1) * sizeof(double);
}

int vb2k_join_slot_flags(const int32_t* head, const int32_t* codes, const uint8_t* flag, int64_t range, uint8_t* out, void* stream) {
  if (range <= 0) return VB2_OK;
  int64_t b = (range + 255) / 256, cap = static_cast<int64_t>(device_sm_count()) * 8;
  join_slot_flags_kernel<<<vb2::counted(static_cast<unsigned>(b > cap ? cap : b)), 256, 0, static_cast<cudaStream_t>(stream)>>>(head, codes, flag, range, out);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_fused_scan_compact(int32_t id, const vb2_fused_args* args, void* const* outs, int32_t nouts, int64_t capacity,
                            int64_t* count, int32_t* error_flag, void* stream) {
  ensure_registered();
  if (id < 0 || id >= static_cast<int>(registry().size()) || !registry()[id].compact) return fail_msg(VB2_ERR_INVALID, "not a scan-compact pipeline");
  const Entry& e = registry()[id];
  if (nouts != e.nproj || !args) return fail_msg(VB2_ERR_INVALID, "scan-compact: wrong number of outputs");
  if (args->rows <= 0) return VB2_OK;
  KernelArgs a;
  std::memset(&a, 0, sizeof(a));
  a.release_guard = kReleaseGuard;
  for (int c = 0; c < kMaxCols; ++c) a.cols[c] = args->cols[c];
  std::memcpy(a.consts.pf, args->pf, sizeof(a.consts.pf));
  std::memcpy(a.consts.pl, args->pl, sizeof(a.consts.pl));
  std::memcpy(a.consts.pi, args->pi, sizeof(a.consts.pi));
  a.rows = args->rows;
  CompactArgs o;
  std::memset(&o, 0, sizeof(o));
  for (int i = 0; i < nouts; ++i) o.outs[i] = outs[i];
  o.capacity = capacity;
  o.count = reinterpret_cast<unsigned long long*>(count);
  o.error_flag = error_flag;
  return e.compact(a, o, static_cast<cudaStream_t>(stream));
}
int32_t vb2k_fused_output_width(int32_t id, int32_t out) {
  ensure_registered();
  if (id < 0 || id >= static_cast<int>(registry().size()) || out < 0 || out >= registry()[id].nproj) return -1;
  return registry()[id].compact ? registry()[id].widths[out] : 8;
}

static int fill_args(const vb2_fused_args* args, KernelArgs& a) {
  if (!args || args->rows < 0 || args->nkeys < 0 || args->nkeys > VB2_FUSED_MAX_KEYS) return fail_msg(VB2_ERR_INVALID, "bad fused args");
  std::memset(&a, 0, sizeof(a));
  a.release_guard = kReleaseGuard;
  for (int c = 0; c < kMaxCols; ++c) a.cols[c] = args->cols[c];
  std::memcpy(a.consts.pf, args->pf, sizeof(a.consts.pf));
  std::memcpy(a.consts.pl, args->pl, sizeof(a.consts.pl));
  std::memcpy(a.consts.pi, args->pi, sizeof(a.consts.pi));
  a.rows = args->rows;
  a.nkeys = args->nkeys;
  a.ngroups = args->nkeys == 0 ? 1 : args->ngroups;
  for (int k = 0; k < VB2_FUSED_MAX_KEYS; ++k) {
    a.key[k] = args->key[k];
    a.key_is64[k] = args->key_is64[k];
    a.key_mult[k] = args->key_mult[k];
    a.key_min[k] = args->key_min[k];
    a.key_lut[k] = args->key_lut[k];
  }
  a.join_slot_flags = args->join_slot_flags;
  a.join_min = args->join_min;
  a.join_range = args->join_range;
  return VB2_OK;
}
static const Entry* entry_of(int32_t id) {
  ensure_registered();
  std::lock_guard<std::mutex> lock(registry_mutex());
  if (id < 0 || id >= static_cast<int>(registry().size())) return nullptr;
  return &registry()[id];
}

int vb2k_fused_scan_agg(int32_t id, const vb2_fused_args* args, double* sums, int64_t* counts, void* workspace,
                        size_t workspace_bytes, void* stream) {
  const Entry* e = entry_of(id);
  if (!e) return fail_msg(VB2_ERR_INVALID, "bad fused kernel id");
  KernelArgs a;
  if (int rc = fill_args(args, a)) return rc;
  if (args->rows == 0) return VB2_OK;
  if (!e->kernels || e->nproj == 0) return fail_msg(VB2_ERR_INVALID, "not an aggregate pipeline");
  if (e->join && (!a.join_slot_flags || a.join_range <= 0)) return fail_msg(VB2_ERR_INVALID, "fused join pipeline needs join_slot_flags");
  return launch_aggregate(*e, a, sums, counts, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int32_t vb2k_fused_has_filter(int32_t id) {
  const Entry* e = entry_of(id);
  return e && e->kernels && e->desc.has_filter ? 1 : 0;
}

int vb2k_fused_filter_bits(int32_t id, const vb2_fused_args* args, int32_t tile_stride, uint64_t* sel_bits, int64_t* counters, void* stream) {
  const Entry* e = entry_of(id);
  if (!e || !e->kernels) return fail_msg(VB2_ERR_INVALID, "bad fused kernel id");
  KernelArgs a;
  if (int rc = fill_args(args, a)) return rc;
  if (args->rows == 0) return VB2_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (sel_bits && tile_stride <= 1) VB2_CUDA_OK(cudaMemsetAsync(sel_bits + ((args->rows + 63) >> 6) - 1, 0, 8, st));  // the last word's unused half
  return launch_filter_bits(*e, a, tile_stride, reinterpret_cast<uint32_t*>(sel_bits), reinterpret_cast<unsigned long long*>(counters), st);
}

int vb2k_fused_gather_agg(int32_t id, const vb2_fused_args* args, const int32_t* sel, const int64_t* nsel_dev, int64_t nsel_hint, double* sums,
                          int64_t* counts, void* workspace, size_t workspace_bytes, void* stream) {
  const Entry* e = entry_of(id);
  if (!e || !e->kernels || e->nproj == 0) return fail_msg(VB2_ERR_INVALID, "bad fused kernel id");
  KernelArgs a;
  if (int rc = fill_args(args, a)) return rc;
  if (args->rows == 0 || !sel || !nsel_dev) return VB2_OK;
  if (e->join && (!a.join_slot_flags || a.join_range <= 0)) return fail_msg(VB2_ERR_INVALID, "fused join pipeline needs join_slot_flags");
  return launch_gather(*e, a, sel, nsel_dev, nsel_hint, sums, counts, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
