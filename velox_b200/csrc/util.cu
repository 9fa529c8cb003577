// Error plumbing and small utility kernels.
#include <atomic>

#include "common.cuh"

#include <cstdio>
#include <string>

namespace vb2 {

static thread_local std::string g_last_error;

int fail_cuda(cudaError_t e, const char* what) {
  g_last_error = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  return VB2_ERR_CUDA;
}
int fail_msg(int code, const char* msg) {
  g_last_error = msg;
  return code;
}
static std::atomic<int64_t> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
int device_sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = kNumSMs;
  }
  return n;
}

__global__ void fill_u64_kernel(uint64_t* p, int64_t n, uint64_t v) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) p[i] = v;
}
__global__ void fill_i32_kernel(int32_t* p, int64_t n, int32_t v) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) p[i] = v;
}

// out bit k = in bit sel[k]
__global__ void gather_bits_kernel(const uint64_t* __restrict__ in, const int32_t* __restrict__ sel, int64_t n, uint32_t* __restrict__ out) {
  const int64_t nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t k = (w << 5) + lane;
    const bool b = k < n && bit_at(in, sel[k]);
    const unsigned word = __ballot_sync(0xffffffffu, b);
    if (lane == 0) out[w] = word;
  }
}
// bit-packed words -> bytes (0/1): validity bitmaps travel through the exchange as byte columns
__global__ void unpack_bits_kernel(const uint64_t* __restrict__ in, int64_t n, uint8_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = bit_at(in, i) ? 1 : 0;
}
// bytes (0/1) -> bit-packed words
__global__ void pack_bools_kernel(const uint8_t* __restrict__ in, int64_t n, uint32_t* __restrict__ out) {
  const int64_t nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t k = (w << 5) + lane;
    const bool b = k < n && in[k] != 0;
    const unsigned word = __ballot_sync(0xffffffffu, b);
    if (lane == 0) out[w] = word;
  }
}
// out[i] = in[i] valid ? ... helpers for aggregation output
__global__ void positive_bits_kernel(const int64_t* __restrict__ counts, int64_t n, uint32_t* __restrict__ out) {
  const int64_t nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t k = (w << 5) + lane;
    const unsigned word = __ballot_sync(0xffffffffu, k < n && counts[k] > 0);
    if (lane == 0) out[w] = word;
  }
}
__global__ void widen_i32_kernel(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out, bool negate) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = negate ? (in[i] == 0 ? 1 : 0) : in[i];
}
// validity bit = idx >= 0; clamped = max(idx, 0)
__global__ void index_validity_kernel(const int32_t* __restrict__ idx, int64_t n, uint32_t* __restrict__ valid, int32_t* __restrict__ clamped) {
  const int64_t nwords = (n + 31) >> 5;
  const int lane = threadIdx.x & 31;
  const int64_t warp_global = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = (static_cast<int64_t>(gridDim.x) * blockDim.x) >> 5;
  for (int64_t w = warp_global; w < nwords; w += nwarps) {
    const int64_t k = (w << 5) + lane;
    bool ok = false;
    if (k < n) {
      const int32_t v = idx[k];
      ok = v >= 0;
      clamped[k] = ok ? v : 0;
    }
    const unsigned word = __ballot_sync(0xffffffffu, ok);
    if (lane == 0) valid[w] = word;
  }
}
__global__ void and_bits_kernel(const uint64_t* __restrict__ a, const uint64_t* __restrict__ b, int64_t nwords, uint64_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < nwords; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[i] = a[i] & (b ? b[i] : ~0ull);
}
__global__ void narrow_i64_kernel(const int64_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) out[i] = static_cast<int32_t>(in[i]);
}
__global__ void iota_i32_kernel(int32_t* p, int64_t n) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x) p[i] = static_cast<int32_t>(i);
}
// out[dst[i]] = in[src ? src[i] : i]
template <class T>
__global__ void scatter_kernel(const T* __restrict__ in, const int32_t* __restrict__ src, const int32_t* __restrict__ dst, int64_t n, T* __restrict__ out) {
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<int64_t>(gridDim.x) * blockDim.x)
    out[dst[i]] = in[src ? src[i] : i];
}

}  // namespace vb2

using namespace vb2;

extern "C" {
const char* vb2_last_error(void) { return g_last_error.c_str(); }
int vb2k_device_sm_count(void) { return device_sm_count(); }
int64_t vb2k_kernel_launches(void) { return g_launches.load(std::memory_order_relaxed); }

static unsigned grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}
int vb2k_fill_u64(uint64_t* p, int64_t n, uint64_t v, void* stream) {
  if (n <= 0) return VB2_OK;
  fill_u64_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, n, v);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_gather_bits(const uint64_t* in, const int32_t* sel, int64_t n, uint64_t* out, void* stream) {
  if (n <= 0) return VB2_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VB2_CUDA_OK(cudaMemsetAsync(out + ((n + 63) >> 6) - 1, 0, 8, st));
  gather_bits_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(in, sel, n, reinterpret_cast<uint32_t*>(out));
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_unpack_bits(const uint64_t* in, int64_t n, uint8_t* out, void* stream) {
  if (n <= 0) return VB2_OK;
  unpack_bits_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, n, out);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_pack_bools(const uint8_t* in, int64_t n, uint64_t* out, void* stream) {
  if (n <= 0) return VB2_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VB2_CUDA_OK(cudaMemsetAsync(out + ((n + 63) >> 6) - 1, 0, 8, st));
  pack_bools_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(in, n, reinterpret_cast<uint32_t*>(out));
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_positive_bits(const int64_t* counts, int64_t n, uint64_t* out, void* stream) {
  if (n <= 0) return VB2_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VB2_CUDA_OK(cudaMemsetAsync(out + ((n + 63) >> 6) - 1, 0, 8, st));
  positive_bits_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(counts, n, reinterpret_cast<uint32_t*>(out));
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_widen_i32(const int32_t* in, int64_t n, int64_t* out, void* stream) {
  if (n <= 0) return VB2_OK;
  widen_i32_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, n, out, false);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_widen_not_i32(const int32_t* in, int64_t n, int64_t* out, void* stream) {
  if (n <= 0) return VB2_OK;
  widen_i32_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, n, out, true);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_index_validity(const int32_t* idx, int64_t n, uint64_t* valid, int32_t* clamped, void* stream) {
  if (n <= 0) return VB2_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  VB2_CUDA_OK(cudaMemsetAsync(valid + ((n + 63) >> 6) - 1, 0, 8, st));
  index_validity_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(idx, n, reinterpret_cast<uint32_t*>(valid), clamped);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_and_bits(const uint64_t* a, const uint64_t* b, int64_t n, uint64_t* out, void* stream) {
  if (n <= 0) return VB2_OK;
  const int64_t nwords = (n + 63) >> 6;
  and_bits_kernel<<<vb2::counted(grid_for(nwords, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(a, b, nwords, out);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_narrow_i64(const int64_t* in, int64_t n, int32_t* out, void* stream) {
  if (n <= 0) return VB2_OK;
  narrow_i64_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(in, n, out);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_iota_i32(int32_t* p, int64_t n, void* stream) {
  if (n <= 0) return VB2_OK;
  iota_i32_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, n);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_scatter(const void* in, const int32_t* src, const int32_t* dst, int64_t n, int32_t elem_bytes, void* out, void* stream) {
  if (n <= 0) return VB2_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (elem_bytes == 8) scatter_kernel<uint64_t><<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(reinterpret_cast<const uint64_t*>(in), src, dst, n, reinterpret_cast<uint64_t*>(out));
  else if (elem_bytes == 4) scatter_kernel<uint32_t><<<vb2::counted(grid_for(n, 256)), 256, 0, st>>>(reinterpret_cast<const uint32_t*>(in), src, dst, n, reinterpret_cast<uint32_t*>(out));
  else return fail_msg(VB2_ERR_INVALID, "scatter: elem_bytes must be 4 or 8");
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_fill_i32(int32_t* p, int64_t n, int32_t v, void* stream) {
  if (n <= 0) return VB2_OK;
  fill_i32_kernel<<<vb2::counted(grid_for(n, 256)), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, n, v);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
}
