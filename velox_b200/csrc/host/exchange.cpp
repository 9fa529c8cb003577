// Hash-partitioned exchange between the GPUs of a node: one grouped ncclSend/ncclRecv all-to-all
// over NVLink. The reference's shuffle is PartitionedOutput -> Exchange over a
// pluggable ExchangeSource (velox/exec/ExchangeSource.h:139-145; GPU variant over UCX in
// velox/experimental/ucx-exchange); partition ids follow HashPartitionFunction
// (velox/exec/HashPartitionFunction.cpp:113-116).
#include <nccl.h>

#include <cstring>
#include <string>
#include <vector>

#include "../../../include/velox_b200.h"
#include "device.h"

struct vb2_comm {
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  velox_b200::DeviceBufferPtr counts;
};

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
  return c;
}

void vb2_comm_free(vb2_comm* comm) {
  if (!comm) return;
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
