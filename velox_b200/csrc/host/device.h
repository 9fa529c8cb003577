// Device-resident vectors and host<->device conversion for the B200 operators.
//
// Pattern of the reference's GPU backends: a RowVector subclass that owns device buffers
// (velox/experimental/cudf/vector/CudfVector.h:43) and a pair of conversion operators inserted
// at CPU/GPU seams (velox/experimental/cudf/exec/CudfConversion.h:32,67).
#pragma once
#include <mutex>
#include <unordered_map>
#include <cuda_runtime.h>

#include <memory>
#include <string>
#include <vector>

#include "../../../include/velox_b200_kernels.h"
#include "../../abi/exec_abi.h"

namespace velox_b200 {

using namespace facebook::velox;

void cudaCheck(cudaError_t e, const char* what);
void kernelCheck(int rc);  // VB2_ERR_USER -> VeloxUserError, others -> VeloxRuntimeError
#define VB2_CU(expr) ::velox_b200::cudaCheck((expr), #expr)

// Stream-ordered allocation: cudaMallocAsync on the driver's stream, freed on the same stream.
class DeviceBuffer {
 public:
  DeviceBuffer(size_t bytes, cudaStream_t stream);
  DeviceBuffer(void* borrowed, size_t bytes) : ptr_(borrowed), bytes_(bytes), stream_(nullptr), owned_(false) {}
  ~DeviceBuffer();
  DeviceBuffer(const DeviceBuffer&) = delete;
  template <class T> T* as() const { return reinterpret_cast<T*>(ptr_); }
  void* data() const { return ptr_; }
  size_t size() const { return bytes_; }

 private:
  void* ptr_ = nullptr;
  size_t bytes_ = 0;
  cudaStream_t stream_;
  bool owned_ = true;
  std::shared_ptr<void> streamOwner_;  // keeps the stream alive until the buffer is freed on it
};
using DeviceBufferPtr = std::shared_ptr<DeviceBuffer>;
DeviceBufferPtr allocDevice(size_t bytes, cudaStream_t stream);
DeviceBufferPtr allocDeviceZeroed(size_t bytes, cudaStream_t stream);

// Host copy of a small dictionary's values (VARCHAR alphabets of flags / types): the planner-side
// metadata needed to assign value ids to group keys and to evaluate build-side predicates.
struct HostAlphabet {
  std::vector<std::string> values;
  std::vector<bool> nulls;
};

struct DeviceColumn {
  TypePtr type;
  vb2_column desc{};                    // device pointers
  std::vector<DeviceBufferPtr> owners;  // buffers referenced by desc
  std::shared_ptr<const HostAlphabet> alphabet;  // DICTIONARY / CONSTANT VARCHAR only
  bool mayHaveNulls() const { return desc.nulls != nullptr || desc.dict_nulls != nullptr; }
};
using DeviceColumnPtr = std::shared_ptr<DeviceColumn>;

// Pinned host memory from a process-wide pool (cudaHostAlloc costs ~100 us; operators that bring
// small results back every query reuse blocks by power-of-two size class).
std::shared_ptr<void> acquirePinned(size_t bytes);

// Host copy of a small device arena: an operator that has to synchronise anyway (e.g. an aggregation
// that needs its output row count) copies its whole packed output to pinned memory in the same
// round trip; B200ToHost then builds the host vectors over this copy without touching the device.
struct HostMirror {
  std::shared_ptr<void> host;   // pinned block
  const uint8_t* devBase = nullptr;
  size_t bytes = 0;
  // host address of a device pointer inside the mirrored arena, nullptr when outside
  const uint8_t* hostOf(const void* dev) const {
    const uint8_t* p = static_cast<const uint8_t*>(dev);
    if (!host || !dev || p < devBase || p >= devBase + bytes) return nullptr;
    return static_cast<const uint8_t*>(host.get()) + (p - devBase);
  }
};

// A batch resident in HBM. children() is empty: the columns live on the device.
class B200Vector : public RowVector {
 public:
  B200Vector(memory::MemoryPool* pool, TypePtr type, vector_size_t size, std::vector<DeviceColumnPtr> cols, cudaStream_t stream)
      : RowVector(pool, std::move(type), nullptr, size, {}), cols_(std::move(cols)), stream_(stream) {}
  const std::vector<DeviceColumnPtr>& columns() const { return cols_; }
  const DeviceColumnPtr& column(size_t i) const { return cols_.at(i); }
  cudaStream_t stream() const { return stream_; }
  // Work on another stream that must finish before the columns may be read (pages that arrive
  // through an exchange are produced on the sending pipeline's stream).
  void setReadyEvent(std::shared_ptr<void> e) { ready_ = std::move(e); }
  cudaEvent_t readyEvent() const { return static_cast<cudaEvent_t>(ready_.get()); }
  void setMirror(std::shared_ptr<const HostMirror> m) { mirror_ = std::move(m); }
  const std::shared_ptr<const HostMirror>& mirror() const { return mirror_; }

 private:
  std::vector<DeviceColumnPtr> cols_;
  cudaStream_t stream_;
  std::shared_ptr<const HostMirror> mirror_;
  std::shared_ptr<void> ready_;
};
using B200VectorPtr = std::shared_ptr<B200Vector>;

// Host RowVector -> device. Flat fixed-width children are one cudaMemcpyAsync each (the vector's
// own buffer is the source: pinned memory gives a true asynchronous DMA); StringView children are
// first rewritten to offsets + chars on the host (pointers cannot be followed by the device).
B200VectorPtr toDevice(const RowVectorPtr& host, cudaStream_t stream);
// Device -> host RowVector (synchronises the stream).
RowVectorPtr toHost(const B200VectorPtr& dev);

// Device copy (int32 offsets + chars) of a small alphabet, cached by content: dictionaries repeat
// from batch to batch and from query to query, the upload happens once per distinct content.
void deviceAlphabet(const HostAlphabet& a, cudaStream_t st, DeviceBufferPtr& offsets, DeviceBufferPtr& chars);

// Wraps externally owned device memory (e.g. columns already resident in HBM) without copying.
DeviceColumnPtr borrowFlatColumn(TypePtr type, const void* values, int64_t size);

// Scan-side device residency for host tables that several tasks read: while a cache is attached to
// the calling thread, host buffers that go through upload() are remembered by (address, bytes) and a
// later upload of the same buffer returns the resident device copy instead of crossing PCIe again.
// The caller guarantees that registered host buffers are neither modified nor freed while the cache
// lives. Uploads happen on the uploading task's stream; every entry carries an event recorded after
// its copy, and a hit from another stream waits on it (tasks may run concurrently, vb2_tasks_run).
struct UploadCache {
  struct Entry {
    size_t bytes = 0;
    DeviceBufferPtr buffer;
    cudaStream_t stream = nullptr;  // stream the copy was issued on
    std::shared_ptr<void> copied;   // cudaEvent_t recorded after the copy
  };
  std::mutex mu;
  std::unordered_map<uint64_t, Entry> entries;  // host address -> device copy
  int64_t hitBytes = 0, missBytes = 0;
};
void setThreadUploadCache(UploadCache* cache);  // nullptr detaches
int64_t threadUploadedBytes();                  // bytes copied host -> device by this thread so far

// Rows [offset, offset + length) of a device batch without copying (offset must be a multiple of 64
// so validity bitmaps stay word-aligned).
B200VectorPtr sliceVector(const B200VectorPtr& v, int64_t offset, int64_t length);

// NVTX range per operator call, named "<Operator>::<method> [planNodeId]" — what the reference's cuDF
// operators push (velox/experimental/cudf/exec/CudfOperator.h:104-160, NvtxHelper.h). NVTX3 is
// header-only: without a profiler attached a push / pop is a load and a branch.
struct NvtxRange {
  NvtxRange(const char* method, const std::string& operatorType, const std::string& planNodeId);
  ~NvtxRange();
  NvtxRange(const NvtxRange&) = delete;
};
#define B200_NVTX_OPERATOR_RANGE(method) ::velox_b200::NvtxRange nvtxRange_(method, stats_.operatorType, planNodeId())

// One stream + small scratch per driver.
struct DeviceContext {
  int device = 0;
  cudaStream_t stream = nullptr;
  DeviceContext();
  ~DeviceContext();
};
std::shared_ptr<DeviceContext> driverDeviceContext(exec::DriverCtx* ctx);

int32_t veloxTypeToVb2(const TypePtr& t);
int32_t widthOf(int32_t vb2Type);

}  // namespace velox_b200
