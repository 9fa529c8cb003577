"""The data path of the GPU-to-GPU shuffle (B200PartitionedOutput -> B200Exchange over peer memory, csrc/exchange_p2p.cu:
p2p_put_rows — the partition gather fused with the stores into the destinations' segments — and p2p_collect) together with
the stable partition order it consumes (csrc/hash_partition.cu), compiled FOR THE HOST and run under the lock-step
emulation of tests/host_emulator.py with W ranks simulated in one process (every rank's exchange heap is a host buffer the
other ranks' kernels store into, as they do through CUDA IPC mappings over NVLink). Checked against numpy: rank r ends up
with exactly the rows whose key hashes to r — HashPartitionFunction, exec/HashPartitionFunction.cpp:75-118 — source by
source, in their original order, every column intact; broadcasts deliver all rows to every rank. No GPU needed: this path
only runs with two or more GPUs, which the single-GPU test box cannot exercise."""
import ctypes as C

import numpy as np
import pytest

from host_emulator import between, build, source
from oracle import pyoracle

BODY = r"""
// ---- common.cuh: hash mixers ----
%(mixers)s
// ---- hash_partition.cu: the stable partition order ----
%(part)s
// ---- exchange_p2p.cu: segment layout, put_rows, collect ----
%(layout)s
%(kernels)s
}  // namespace vb2_on_host
using namespace vb2_on_host;
extern "C" {
int64_t h_segment_bytes(const int32_t* widths, int ncols, int64_t rows) { return p2p_col_offset(widths, ncols, rows); }
// one source rank: partition its rows by key and store them into every destination's segment for this source
void h_send(const int64_t* keys, int64_t n, int world, const void* const* cols, const int32_t* widths, int ncols, void* const* dest_segments,
            int64_t* counts_out, int broadcast) {
  std::vector<int32_t> order(n > 0 ? n : 1);
  std::vector<int64_t> counts(world, 0);
  if (!broadcast && n > 0) {
    const PartSrc src{nullptr, keys, 1};
    const int64_t nblocks = (n + kPartRowsPerBlock - 1) / kPartRowsPerBlock;
    std::vector<int32_t> hist(nblocks * world);
    std::vector<int64_t> base(nblocks * world);
    launch(static_cast<unsigned>(nblocks), kPartThreads, [&] { part_hist_kernel(src, n, world, hist.data()); });
    launch(1, kOffsetThreads, [&] { part_offsets_kernel(hist.data(), nblocks, world, counts.data(), base.data()); });
    launch(static_cast<unsigned>(nblocks), kPartThreads, [&] { part_scatter_kernel(src, n, world, base.data(), order.data()); });
  }
  for (int p = 0; p < world; ++p) counts_out[p] = broadcast ? n : counts[p];
  P2pCols pc{};
  pc.n = ncols;
  for (int c = 0; c < ncols; ++c) { pc.src[c] = cols[c]; pc.width[c] = widths[c]; }
  P2pPeers seg{};
  for (int p = 0; p < world; ++p) seg.p[p] = dest_segments[p];
  if (n > 0) launch(2, 256, [&] { p2p_put_rows_kernel(broadcast ? nullptr : order.data(), counts.data(), world, n, pc, seg, broadcast); });
}
// one destination rank: its W source segments -> contiguous output columns
void h_collect(const void* const* my_segments, const int64_t* recv_counts, int world, const int32_t* widths, int ncols, void* const* outs) {
  P2pCollect a{};
  a.world = world;
  a.ncols = ncols;
  int64_t run = 0;
  for (int s = 0; s < world; ++s) { a.seg[s] = my_segments[s]; a.count[s] = recv_counts[s]; a.row_start[s] = run; run += recv_counts[s]; }
  for (int c = 0; c < ncols; ++c) { a.out[c] = outs[c]; a.width[c] = widths[c]; }
  for (int s = 0; s < world; ++s)
    for (int c = 0; c < ncols; ++c) {
      blockIdx.y = s;
      blockIdx.z = c;
      launch(1, 256, [&] { p2p_collect_kernel(a); });
    }
  blockIdx.y = blockIdx.z = 0;
}
}
"""


@pytest.fixture(scope="module")
def host(tmp_path_factory):
    common, part, p2p = source("common.cuh"), source("hash_partition.cu"), source("exchange_p2p.cu")
    body = BODY % {
        "mixers": between(common, "__host__ __device__ __forceinline__ uint64_t twang_mix64", "__device__ __forceinline__ uint64_t hash_f64"),
        "part": between(part, "constexpr int kPartThreads", "// --- fixed-capacity segments"),
        "layout": between(p2p, "constexpr int kP2pMaxWorld", "__device__ __forceinline__ void st_release_sys"),
        "kernels": between(p2p, "// Rows grouped by destination", "}  // namespace vb2"),
    }
    L = build(tmp_path_factory.mktemp("exchange_on_host"), "exchange", body)
    L.h_segment_bytes.restype = C.c_int64
    return L


def _ptrs(arrays):
    return (C.c_void_p * len(arrays))(*[a.ctypes.data for a in arrays])


@pytest.mark.parametrize("world", [2, 5])
@pytest.mark.parametrize("broadcast", [0, 1])
def test_shuffle_between_simulated_ranks(host, world, broadcast):
    rng = np.random.default_rng(world * 2 + broadcast)
    widths = np.array([8, 8, 4, 1], dtype=np.int32)  # key, DOUBLE payload, int32 dictionary codes, validity bytes
    rows = [int(v) for v in rng.integers(0, 3000, world)]
    rows[0] = 4096 + 517                               # one rank with several partition blocks
    rows[-1] = 0                                       # and one with nothing to send
    data = []
    for s in range(world):
        n = rows[s]
        data.append([rng.integers(-2**60, 2**60, n), np.round(rng.normal(size=n), 3), rng.integers(0, 150, n).astype(np.int32),
                     (rng.random(n) < 0.9).astype(np.uint8)])
    cap = max(rows) if broadcast else max(rows)
    seg_bytes = int(host.h_segment_bytes(widths.ctypes.data_as(C.c_void_p), 4, C.c_int64(cap))) + 256
    # heap[r][s]: the segment of rank r that source s writes (the exchange heap every peer maps)
    heap = [[np.zeros(seg_bytes, dtype=np.uint8) for _ in range(world)] for _ in range(world)]
    sent = np.zeros((world, world), dtype=np.int64)
    for s in range(world):
        dest = _ptrs([heap[r][s] for r in range(world)])
        counts = np.zeros(world, dtype=np.int64)
        host.h_send(data[s][0].ctypes.data_as(C.c_void_p), C.c_int64(rows[s]), world, _ptrs(data[s]), widths.ctypes.data_as(C.c_void_p), 4, dest,
                    counts.ctypes.data_as(C.c_void_p), broadcast)
        sent[s] = counts
    L = pyoracle.lib()
    for r in range(world):
        recv = np.ascontiguousarray(sent[:, r])
        total = int(recv.sum())
        outs = [np.zeros(total + 1, dtype=dt) for dt in (np.int64, np.float64, np.int32, np.uint8)]
        host.h_collect(_ptrs(heap[r]), recv.ctypes.data_as(C.c_void_p), world, widths.ctypes.data_as(C.c_void_p), 4, _ptrs(outs))
        want = [[], [], [], []]
        for s in range(world):
            if broadcast:
                keep = np.ones(rows[s], dtype=bool)
            else:
                ids = np.array([L.orc_twang_mix64(int(k) & 0xFFFFFFFFFFFFFFFF) % world for k in data[s][0]], dtype=np.int64)
                keep = ids == r
            assert int(recv[s]) == int(keep.sum())
            for c in range(4):
                want[c].append(data[s][c][keep])
        for c in range(4):
            assert np.array_equal(outs[c][:total], np.concatenate(want[c])), (r, c)
    assert int(sent.sum()) == (sum(rows) * world if broadcast else sum(rows))  # conservation: rows sent = rows received
