// Hash-partitioned exchange between the GPUs of a node: one grouped ncclSend/ncclRecv all-to-all
// over NVLink. The reference's shuffle is PartitionedOutput -> Exchange over a
// pluggable ExchangeSource (velox/exec/ExchangeSource.h:139-145; GPU variant over UCX in
// velox/experimental/ucx-exchange); partition ids follow HashPartitionFunction
// (velox/exec/HashPartitionFunction.cpp:113-116).
#include <nccl.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../../include/velox_b200.h"
#include "device.h"

#include "exchange_impl.h"

struct vb2_comm {
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  velox_b200::DeviceBufferPtr counts;
  // ---- peer-memory exchange (exchange_p2p.cu): every rank's heap mapped into every process ----
  bool p2p = false;
  uint8_t* heap = nullptr;            // this rank's heap
  std::vector<uint8_t*> peer;         // [world], peer[rank] == heap
  size_t segBytes = 0;                // capacity of one (parity, source) data segment
  size_t metaOff = 0, dataOff = 0, heapBytes = 0;
  cudaStream_t xstream = nullptr;     // every protocol step of this process, in order
  uint64_t metaEpoch = 0, dataEpoch = 0;
  int32_t* errFlag = nullptr;         // device: set by a wait that timed out
  cudaEvent_t orderEvent = nullptr;   // reused: orders the exchange stream behind the producing stream
  int64_t p2pExchanges = 0, ncclExchanges = 0;
};

namespace {
constexpr size_t kFlagStride = 128;                 // bytes between flag words
constexpr size_t kMetaBytes = 64 * 1024;            // metadata block per source
constexpr uint64_t kWaitTimeoutNs = 20ull * 1000 * 1000 * 1000;

uint8_t* metaFlagOf(vb2_comm* c, int owner, int src) { return c->peer[owner] + static_cast<size_t>(src) * kFlagStride; }
uint8_t* dataFlagOf(vb2_comm* c, int owner, int src) { return c->peer[owner] + (static_cast<size_t>(c->world) + src) * kFlagStride; }
uint8_t* metaBlockOf(vb2_comm* c, int owner, int src) { return c->peer[owner] + c->metaOff + static_cast<size_t>(src) * kMetaBytes; }
uint8_t* segmentOf(vb2_comm* c, int owner, int src, uint64_t epoch) {
  return c->peer[owner] + c->dataOff + ((epoch & 1) * c->world + src) * c->segBytes;
}

// Allocates the heap, trades CUDA IPC handles over NCCL and maps every peer. All ranks end with the
// same answer: peer-memory exchange on, or (any failure anywhere) NCCL only.
void setupPeerHeap(vb2_comm* c) {
  const char* mode = std::getenv("VB2_EXCHANGE");
  if (mode && std::string(mode) == "nccl") return;
  if (c->world < 2 || c->world > 16) return;
  const char* segEnv = std::getenv("VB2_EXCHANGE_SEGMENT_MB");
  c->segBytes = static_cast<size_t>(segEnv ? std::atoll(segEnv) : 64) << 20;
  c->metaOff = (2 * static_cast<size_t>(c->world) * kFlagStride + 4095) / 4096 * 4096;
  c->dataOff = c->metaOff + static_cast<size_t>(c->world) * kMetaBytes;
  c->heapBytes = c->dataOff + 2 * static_cast<size_t>(c->world) * c->segBytes;
  int32_t ok = 1;
  cudaIpcMemHandle_t mine{};
  if (cudaStreamCreateWithFlags(&c->xstream, cudaStreamNonBlocking) != cudaSuccess) ok = 0;
  if (ok && cudaMalloc(reinterpret_cast<void**>(&c->heap), c->heapBytes) != cudaSuccess) { ok = 0; c->heap = nullptr; }
  if (ok && cudaMemset(c->heap, 0, c->dataOff) != cudaSuccess) ok = 0;
  if (ok && cudaMalloc(reinterpret_cast<void**>(&c->errFlag), 8) != cudaSuccess) ok = 0;
  if (ok) cudaMemset(c->errFlag, 0, 8);
  if (ok && cudaIpcGetMemHandle(&mine, c->heap) != cudaSuccess) ok = 0;
  cudaGetLastError();
  // handles + success flags travel over NCCL (plumbing, once per communicator)
  const size_t hb = sizeof(cudaIpcMemHandle_t);
  uint8_t* dbuf = nullptr;
  if (cudaMalloc(reinterpret_cast<void**>(&dbuf), (hb + 8) * (c->world + 1)) != cudaSuccess) { cudaGetLastError(); return; }
  std::vector<uint8_t> sendBlock(hb + 8, 0), all((hb + 8) * c->world, 0);
  std::memcpy(sendBlock.data(), &mine, hb);
  std::memcpy(sendBlock.data() + hb, &ok, 4);
  cudaMemcpy(dbuf, sendBlock.data(), hb + 8, cudaMemcpyHostToDevice);
  cudaStream_t st = c->xstream ? c->xstream : nullptr;
  if (ncclAllGather(dbuf, dbuf + (hb + 8), hb + 8, ncclUint8, c->comm, st) != ncclSuccess) { cudaFree(dbuf); return; }
  cudaStreamSynchronize(st);
  cudaMemcpy(all.data(), dbuf + (hb + 8), (hb + 8) * c->world, cudaMemcpyDeviceToHost);
  int32_t everyone = 1;
  for (int p = 0; p < c->world; ++p) {
    int32_t f;
    std::memcpy(&f, all.data() + p * (hb + 8) + hb, 4);
    everyone = everyone && f;
  }
  c->peer.assign(c->world, nullptr);
  int32_t mapped = everyone;
  if (everyone) {
    for (int p = 0; p < c->world && mapped; ++p) {
      if (p == c->rank) { c->peer[p] = c->heap; continue; }
      cudaIpcMemHandle_t h;
      std::memcpy(&h, all.data() + p * (hb + 8), hb);
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { mapped = 0; cudaGetLastError(); }
      c->peer[p] = static_cast<uint8_t*>(ptr);
    }
  }
  // second round: did every rank map every peer?
  cudaMemcpy(dbuf, &mapped, 4, cudaMemcpyHostToDevice);
  if (ncclAllReduce(dbuf, dbuf, 1, ncclInt32, ncclMin, c->comm, st) != ncclSuccess) { cudaFree(dbuf); return; }
  cudaStreamSynchronize(st);
  int32_t agreed = 0;
  cudaMemcpy(&agreed, dbuf, 4, cudaMemcpyDeviceToHost);
  cudaFree(dbuf);
  c->p2p = agreed != 0;
}
}  // namespace

namespace velox_b200 {

bool exchangeUsesPeerMemory(vb2_comm* c) { return c && c->p2p; }
size_t exchangeMaxMetadataBytes(vb2_comm* c) { return c && c->p2p ? kMetaBytes : (1u << 20); }

std::shared_ptr<void> exchangeMetadata(vb2_comm* c, const uint8_t* myBlockHost, size_t blockBytes, const int64_t* countsDev,
                                       const std::vector<ExchangePatch>& patches, cudaStream_t after) {
  const int w = c->world;
  VELOX_CHECK(blockBytes % 16 == 0, "exchange metadata block must be a multiple of 16 bytes");
  cudaStream_t xs = c->p2p ? c->xstream : after;
  auto staging = acquirePinned(blockBytes);
  std::memcpy(staging.get(), myBlockHost, blockBytes);
  auto blockDev = allocDevice(blockBytes, xs);
  auto allHost = acquirePinned(blockBytes * w + 16);
  if (c->p2p && (countsDev || !patches.empty())) {
    // device-side inputs of the block were produced on the caller's stream
    if (!c->orderEvent) VB2_CU(cudaEventCreateWithFlags(&c->orderEvent, cudaEventDisableTiming));
    VB2_CU(cudaEventRecord(c->orderEvent, after));
    VB2_CU(cudaStreamWaitEvent(xs, c->orderEvent, 0));
  }
  VB2_CU(cudaMemcpyAsync(blockDev->data(), staging.get(), blockBytes, cudaMemcpyHostToDevice, xs));
  if (countsDev) VB2_CU(cudaMemcpyAsync(blockDev->data(), countsDev, static_cast<size_t>(w) * 8, cudaMemcpyDeviceToDevice, xs));
  for (auto& pt : patches) VB2_CU(cudaMemcpyAsync(blockDev->as<uint8_t>() + pt.offset, pt.src, pt.bytes, cudaMemcpyDeviceToDevice, xs));
  if (c->p2p) {
    VELOX_CHECK(blockBytes <= kMetaBytes, "exchange metadata block above the peer-memory limit");
    const uint64_t epoch = ++c->metaEpoch;
    std::vector<void*> dst(w), flags(w);
    for (int p = 0; p < w; ++p) { dst[p] = metaBlockOf(c, p, c->rank); flags[p] = metaFlagOf(c, p, c->rank); }
    kernelCheck(vb2k_p2p_put_block(dst.data(), w, blockDev->data(), static_cast<int64_t>(blockBytes), xs));
    kernelCheck(vb2k_p2p_signal(flags.data(), w, epoch, xs));
    kernelCheck(vb2k_p2p_wait(reinterpret_cast<const uint64_t*>(metaFlagOf(c, c->rank, 0)), kFlagStride / 8, w, epoch, c->errFlag, kWaitTimeoutNs, xs));
    // the W blocks lie kMetaBytes apart in the heap: one strided copy
    VB2_CU(cudaMemcpy2DAsync(allHost.get(), blockBytes, metaBlockOf(c, c->rank, 0), kMetaBytes, blockBytes, w, cudaMemcpyDeviceToHost, xs));
    VB2_CU(cudaMemcpyAsync(static_cast<uint8_t*>(allHost.get()) + blockBytes * w, c->errFlag, 4, cudaMemcpyDeviceToHost, xs));
  } else {
    auto allDev = allocDevice(blockBytes * w, xs);
    VELOX_CHECK(ncclAllGather(blockDev->data(), allDev->data(), blockBytes, ncclUint8, c->comm, xs) == ncclSuccess, "exchange metadata all-gather failed");
    VB2_CU(cudaMemcpyAsync(allHost.get(), allDev->data(), blockBytes * w, cudaMemcpyDeviceToHost, xs));
    std::memset(static_cast<uint8_t*>(allHost.get()) + blockBytes * w, 0, 4);
  }
  VB2_CU(cudaStreamSynchronize(xs));  // the exchange's only host synchronisation
  int32_t err;
  std::memcpy(&err, static_cast<uint8_t*>(allHost.get()) + blockBytes * w, 4);
  if (err != 0) throw VeloxRuntimeError("exchange: rank " + std::to_string(err - 200) + " did not arrive (peer-memory wait timed out)");
  return allHost;
}

std::shared_ptr<void> exchangePayload(vb2_comm* c, const int32_t* order, const int64_t* countsDev, const int64_t* matrix, int64_t n,
                                      const std::vector<const void*>& cols, const std::vector<int32_t>& widths, const std::vector<void*>& outs,
                                      bool broadcast, cudaStream_t after, bool* usedPeerMemory) {
  const int w = c->world, me = c->rank;
  const int ncols = static_cast<int>(cols.size());
  std::vector<int64_t> sendCounts(w), recvCounts(w);
  for (int p = 0; p < w; ++p) { sendCounts[p] = matrix[me * w + p]; recvCounts[p] = matrix[p * w + me]; }
  // one decision for all ranks, from the matrix every rank holds: does every (source, destination) block fit a segment?
  bool fits = c->p2p && ncols <= 24;
  if (fits)
    for (int i = 0; i < w * w; ++i) fits = fits && static_cast<size_t>(vb2k_p2p_segment_bytes(widths.data(), ncols, matrix[i])) <= c->segBytes;
  if (usedPeerMemory) *usedPeerMemory = fits;
  cudaStream_t xs = c->p2p ? c->xstream : after;
  cudaEvent_t ev;
  VB2_CU(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
  if (c->p2p) {
    VB2_CU(cudaEventRecord(ev, after));
    VB2_CU(cudaStreamWaitEvent(xs, ev, 0));
  }
  if (fits) {
    ++c->p2pExchanges;
    const uint64_t epoch = ++c->dataEpoch;
    std::vector<void*> seg(w), flags(w);
    for (int p = 0; p < w; ++p) { seg[p] = segmentOf(c, p, me, epoch); flags[p] = dataFlagOf(c, p, me); }
    // gather + transfer in one kernel: rows go from the source columns straight into the peers' heaps
    kernelCheck(vb2k_p2p_put_rows(order, countsDev, w, n, cols.data(), widths.data(), ncols, seg.data(), broadcast ? 1 : 0, xs));
    kernelCheck(vb2k_p2p_signal(flags.data(), w, epoch, xs));
    kernelCheck(vb2k_p2p_wait(reinterpret_cast<const uint64_t*>(dataFlagOf(c, me, 0)), kFlagStride / 8, w, epoch, c->errFlag, kWaitTimeoutNs, xs));
    std::vector<const void*> local(w);
    for (int s = 0; s < w; ++s) local[s] = segmentOf(c, me, s, epoch);
    kernelCheck(vb2k_p2p_collect(local.data(), recvCounts.data(), w, widths.data(), ncols, outs.data(), xs));
  } else {
    ++c->ncclExchanges;
    // NCCL: group the rows by destination first, then one grouped send/recv for all columns
    std::vector<DeviceBufferPtr> keep;
    std::vector<const void*> send(ncols);
    for (int i = 0; i < ncols; ++i) {
      send[i] = cols[i];
      if (order && n > 0) {
        auto g = allocDevice(static_cast<size_t>(n) * widths[i], xs);
        kernelCheck(vb2k_gather(cols[i], order, n, widths[i], g->data(), xs));
        keep.push_back(g);
        send[i] = g->data();
      }
    }
    int rc;
    if (broadcast) rc = vb2_comm_all_gather_columns(c, ncols, send.data(), outs.data(), widths.data(), n, recvCounts.data(), xs);
    else rc = vb2_comm_all_to_all_columns(c, ncols, send.data(), outs.data(), widths.data(), sendCounts.data(), recvCounts.data(), xs);
    VELOX_CHECK(rc == VB2_OK, "exchange all-to-all failed");
  }
  VB2_CU(cudaEventRecord(ev, xs));
  return std::shared_ptr<void>(ev, [](void* p) { cudaEventDestroy(static_cast<cudaEvent_t>(p)); });
}

}  // namespace velox_b200

namespace {
void setErr(char* err, int32_t errlen, const std::string& msg) {
  if (!err || errlen <= 0) return;
  std::strncpy(err, msg.c_str(), errlen - 1);
  err[errlen - 1] = 0;
}
int ncclFail(ncclResult_t r, const char* what) {
  (void)what;
  return r == ncclSuccess ? VB2_OK : VB2_ERR_CUDA;
}
}  // namespace

extern "C" {

int32_t vb2_comm_unique_id(uint8_t out[128]) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  if (ncclGetUniqueId(&id) != ncclSuccess) return VB2_ERR_CUDA;
  std::memcpy(out, &id, 128);
  return VB2_OK;
}

vb2_comm* vb2_comm_create(const uint8_t unique_id[128], int32_t world, int32_t rank, char* err, int32_t errlen) {
  auto c = new vb2_comm();
  c->world = world;
  c->rank = rank;
  ncclUniqueId id;
  std::memcpy(&id, unique_id, 128);
  const ncclResult_t r = ncclCommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) {
    setErr(err, errlen, std::string("ncclCommInitRank: ") + ncclGetErrorString(r));
    delete c;
    return nullptr;
  }
  setupPeerHeap(c);
  return c;
}

int32_t vb2_comm_peer_memory(vb2_comm* comm) { return comm && comm->p2p ? 1 : 0; }
int64_t vb2_comm_exchanges(vb2_comm* comm, int32_t peer_memory) { return !comm ? 0 : (peer_memory ? comm->p2pExchanges : comm->ncclExchanges); }

void vb2_comm_free(vb2_comm* comm) {
  if (!comm) return;
  if (comm->xstream) cudaStreamSynchronize(comm->xstream);
  for (int p = 0; p < static_cast<int>(comm->peer.size()); ++p)
    if (p != comm->rank && comm->peer[p]) cudaIpcCloseMemHandle(comm->peer[p]);
  if (comm->heap) cudaFree(comm->heap);
  if (comm->errFlag) cudaFree(comm->errFlag);
  if (comm->orderEvent) cudaEventDestroy(comm->orderEvent);
  if (comm->xstream) cudaStreamDestroy(comm->xstream);
  if (comm->comm) ncclCommDestroy(comm->comm);
  delete comm;
}

int32_t vb2_comm_exchange_counts(vb2_comm* comm, const int64_t* send_counts, int64_t* recv_counts, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int w = comm->world;
  try {
    auto buf = velox_b200::allocDevice(static_cast<size_t>(w) * 16, st);
    int64_t* dsend = buf->as<int64_t>();
    int64_t* drecv = dsend + w;
    VB2_CU(cudaMemcpyAsync(dsend, send_counts, w * 8, cudaMemcpyHostToDevice, st));
    ncclGroupStart();
    for (int p = 0; p < w; ++p) {
      ncclSend(dsend + p, 1, ncclInt64, p, comm->comm, st);
      ncclRecv(drecv + p, 1, ncclInt64, p, comm->comm, st);
    }
    if (ncclGroupEnd() != ncclSuccess) return VB2_ERR_CUDA;
    VB2_CU(cudaMemcpyAsync(recv_counts, drecv, w * 8, cudaMemcpyDeviceToHost, st));
    VB2_CU(cudaStreamSynchronize(st));
  } catch (const std::exception&) {
    return VB2_ERR_CUDA;
  }
  return VB2_OK;
}

int32_t vb2_comm_exchange_counts_dev(vb2_comm* comm, const int64_t* dev_send_counts, int64_t* send_counts_host, int64_t* recv_counts_host,
                                     void* stream) {
  // counts stay on the device for the exchange; one device->host copy (and one sync) brings both
  // the send and the receive counts back for sizing the payload all-to-all
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int w = comm->world;
  try {
    auto buf = velox_b200::allocDevice(static_cast<size_t>(w) * 16, st);  // per call: exchanges may overlap on two streams
    int64_t* both = buf->as<int64_t>();
    VB2_CU(cudaMemcpyAsync(both, dev_send_counts, w * 8, cudaMemcpyDeviceToDevice, st));
    ncclGroupStart();
    for (int p = 0; p < w; ++p) {
      ncclSend(both + p, 1, ncclInt64, p, comm->comm, st);
      ncclRecv(both + w + p, 1, ncclInt64, p, comm->comm, st);
    }
    if (ncclGroupEnd() != ncclSuccess) return VB2_ERR_CUDA;
    std::vector<int64_t> h(static_cast<size_t>(w) * 2);
    VB2_CU(cudaMemcpyAsync(h.data(), both, w * 16, cudaMemcpyDeviceToHost, st));
    VB2_CU(cudaStreamSynchronize(st));
    std::memcpy(send_counts_host, h.data(), w * 8);
    std::memcpy(recv_counts_host, h.data() + w, w * 8);
  } catch (const std::exception&) {
    return VB2_ERR_CUDA;
  }
  return VB2_OK;
}

int32_t vb2_comm_all_to_all(vb2_comm* comm, const void* send, const int64_t* send_counts, void* recv, const int64_t* recv_counts,
                            int32_t elem_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int w = comm->world;
  const char* s = static_cast<const char*>(send);
  char* r = static_cast<char*>(recv);
  int64_t soff = 0, roff = 0;
  ncclGroupStart();
  for (int p = 0; p < w; ++p) {
    if (send_counts[p] > 0) ncclSend(s + soff * elem_bytes, static_cast<size_t>(send_counts[p]) * elem_bytes, ncclUint8, p, comm->comm, st);
    if (recv_counts[p] > 0) ncclRecv(r + roff * elem_bytes, static_cast<size_t>(recv_counts[p]) * elem_bytes, ncclUint8, p, comm->comm, st);
    soff += send_counts[p];
    roff += recv_counts[p];
  }
  return ncclFail(ncclGroupEnd(), "all_to_all");
}

int32_t vb2_comm_all_to_all_columns(vb2_comm* comm, int32_t ncols, const void* const* send, void* const* recv, const int32_t* elem_bytes,
                                    const int64_t* send_counts, const int64_t* recv_counts, void* stream) {
  // every column of the row set moves inside ONE NCCL group (one fused send/recv kernel)
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int w = comm->world;
  ncclGroupStart();
  for (int c = 0; c < ncols; ++c) {
    const char* s = static_cast<const char*>(send[c]);
    char* r = static_cast<char*>(recv[c]);
    int64_t soff = 0, roff = 0;
    for (int p = 0; p < w; ++p) {
      if (send_counts[p] > 0) ncclSend(s + soff * elem_bytes[c], static_cast<size_t>(send_counts[p]) * elem_bytes[c], ncclUint8, p, comm->comm, st);
      if (recv_counts[p] > 0) ncclRecv(r + roff * elem_bytes[c], static_cast<size_t>(recv_counts[p]) * elem_bytes[c], ncclUint8, p, comm->comm, st);
      soff += send_counts[p];
      roff += recv_counts[p];
    }
  }
  return ncclFail(ncclGroupEnd(), "all_to_all_columns");
}

int32_t vb2_comm_all_gather(vb2_comm* comm, const void* send, void* recv, int64_t bytes, void* stream) {
  return ncclFail(ncclAllGather(send, recv, static_cast<size_t>(bytes), ncclUint8, comm->comm, static_cast<cudaStream_t>(stream)), "all_gather");
}

int32_t vb2_comm_all_gather_columns(vb2_comm* comm, int32_t ncols, const void* const* send, void* const* recv, const int32_t* elem_bytes,
                                    int64_t send_rows, const int64_t* recv_counts, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int w = comm->world;
  ncclGroupStart();
  for (int c = 0; c < ncols; ++c) {
    char* r = static_cast<char*>(recv[c]);
    int64_t roff = 0;
    for (int p = 0; p < w; ++p) {
      if (send_rows > 0) ncclSend(send[c], static_cast<size_t>(send_rows) * elem_bytes[c], ncclUint8, p, comm->comm, st);
      if (recv_counts[p] > 0) ncclRecv(r + roff * elem_bytes[c], static_cast<size_t>(recv_counts[p]) * elem_bytes[c], ncclUint8, p, comm->comm, st);
      roff += recv_counts[p];
    }
  }
  return ncclFail(ncclGroupEnd(), "all_gather_columns");
}

int32_t vb2_comm_world(vb2_comm* comm) { return comm ? comm->world : 1; }
int32_t vb2_comm_rank(vb2_comm* comm) { return comm ? comm->rank : 0; }

int32_t vb2_comm_all_reduce_f64(vb2_comm* comm, double* data, int64_t n, void* stream) {
  return ncclFail(ncclAllReduce(data, data, static_cast<size_t>(n), ncclDouble, ncclSum, comm->comm, static_cast<cudaStream_t>(stream)), "all_reduce");
}
int32_t vb2_comm_all_reduce_i64(vb2_comm* comm, int64_t* data, int64_t n, void* stream) {
  return ncclFail(ncclAllReduce(data, data, static_cast<size_t>(n), ncclInt64, ncclSum, comm->comm, static_cast<cudaStream_t>(stream)), "all_reduce");
}
int32_t vb2_comm_all_reduce_max_i64(vb2_comm* comm, int64_t* data, int64_t n, void* stream) {
  return ncclFail(ncclAllReduce(data, data, static_cast<size_t>(n), ncclInt64, ncclMax, comm->comm, static_cast<cudaStream_t>(stream)), "all_reduce_max");
}

}  // extern "C"
