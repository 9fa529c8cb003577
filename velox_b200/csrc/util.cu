// Error plumbing and small utility kernels.
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

}  // namespace vb2

using namespace vb2;

extern "C" {
const char* vb2_last_error(void) { return g_last_error.c_str(); }
int vb2k_device_sm_count(void) { return device_sm_count(); }

static unsigned grid_for(int64_t n, int threads) {
  int64_t b = (n + threads - 1) / threads;
  int64_t cap = static_cast<int64_t>(device_sm_count()) * 16;
  return static_cast<unsigned>(b < 1 ? 1 : (b > cap ? cap : b));
}
int vb2k_fill_u64(uint64_t* p, int64_t n, uint64_t v, void* stream) {
  if (n <= 0) return VB2_OK;
  fill_u64_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, n, v);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
int vb2k_fill_i32(int32_t* p, int64_t n, int32_t v, void* stream) {
  if (n <= 0) return VB2_OK;
  fill_i32_kernel<<<grid_for(n, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(p, n, v);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}
}
