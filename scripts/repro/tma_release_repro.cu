// Minimal repro for the stage-release rule of the TMA pipelines (velox_b200/csrc/fused_scan.cuh):
// a producer warp streams 1024-row int64 tiles into a 2-stage shared-memory ring with 1-D bulk
// copies; 256 consumer threads read 4 values each, RELEASE the stage, and only then use the values.
// The column holds row numbers, so the exact sum is n(n-1)/2: any stale or torn tile shows up as a
// wrong integer. MODE selects how the stage is released:
//   0  lane 0 arrives after __syncwarp(), no data dependency on the loads        (count = 8)
//   1  every consumer thread arrives for itself, no data dependency              (count = 256)
//   2  lane 0 arrives, predicated on a fold of its loaded registers (shipping)   (count = 8)
//   3  mode 0 + fence.proxy.async.shared::cta before the arrive (consumer side)
//   4  mode 0 + fence.proxy.async.shared::cta in the producer after its wait
//   5  every lane folds its loads, __reduce-style xor across the warp feeds lane 0's predicate
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo tma_release_repro.cu -o tma_release_repro
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cuda_runtime.h>
constexpr int kTile = 1024, kConsumers = 256, kStages = 2;
__device__ __forceinline__ uint32_t s32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(s32(b)), "r"(parity) : "memory");
}
template <int MODE>
__global__ void __launch_bounds__(kConsumers + 32) repro(const int64_t* __restrict__ col, int64_t rows, unsigned long long* out, uint64_t guard) {
  __shared__ __align__(128) int64_t tile[kStages][kTile];
  __shared__ uint64_t full[kStages], empty[kStages];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < kStages; ++s) {
      asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&full[s])));
      asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(&empty[s])), "r"(MODE == 1 ? kConsumers : kConsumers / 32));
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int64_t ntiles = rows / kTile;
  if (warp == kConsumers / 32) {
    if (lane == 0) {
      int it = 0;
      for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
        const int s = it % kStages;
        if (it >= kStages) mbar_wait(&empty[s], ((it / kStages) - 1) & 1);
        if (MODE == 4) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&full[s])), "r"(kTile * 8) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(tile[s])), "l"(col + t * kTile), "r"(kTile * 8), "r"(s32(&full[s])) : "memory");
      }
    }
    return;
  }
  unsigned long long acc = 0;
  int it = 0;
  for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x, ++it) {
    const int s = it % kStages;
    mbar_wait(&full[s], (it / kStages) & 1);
    int64_t v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = tile[s][j * kConsumers + threadIdx.x];
    if (MODE == 1) {
      asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
    } else {
      uint64_t dep = static_cast<uint64_t>(v[0] ^ v[1] ^ v[2] ^ v[3]);
      if (MODE == 5) for (int o = 16; o > 0; o >>= 1) dep ^= __shfl_xor_sync(0xffffffffu, dep, o);
      if (MODE == 3) asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (MODE == 2 || MODE == 5)
          asm volatile("{ .reg .pred q; setp.ne.b64 q, %1, %2; @q mbarrier.arrive.shared::cta.b64 _, [%0]; @!q mbarrier.arrive.shared::cta.b64 _, [%0], 1; }" ::"r"(s32(&empty[s])), "l"(dep), "l"(guard) : "memory");
        else
          asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(&empty[s])) : "memory");
      }
    }
    acc += static_cast<unsigned long long>(v[0]) + v[1] + v[2] + v[3];  // values are consumed AFTER the release
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) atomicAdd(out, acc);
}
__global__ void fill(int64_t* p, int64_t n) {
  for (int64_t i = blockIdx.x * 256ll + threadIdx.x; i < n; i += gridDim.x * 256ll) p[i] = i;
}
template <int MODE>
int run(const int64_t* col, int64_t rows, unsigned long long* out, int iters) {
  const unsigned long long want = static_cast<unsigned long long>(rows) * (rows - 1) / 2;
  int bad = 0;
  for (int i = 0; i < iters; ++i) {
    cudaMemset(out, 0, 8);
    repro<MODE><<<148 * 2, kConsumers + 32>>>(col, rows, out, 0x9e3779b97f4a7c15ull);
    unsigned long long got = 0;
    cudaMemcpy(&got, out, 8, cudaMemcpyDeviceToHost);
    bad += got != want;
  }
  cudaError_t e = cudaGetLastError();
  printf("{\"mode\": %d, \"runs\": %d, \"wrong_sums\": %d, \"cuda\": \"%s\"}\n", MODE, iters, bad, cudaGetErrorString(e));
  return bad;
}
int main(int argc, char** argv) {
  const int64_t rows = 1024ll * (argc > 1 ? atoll(argv[1]) : 256 * 1024);  // default 256 Mi rows = 2 GiB
  const int iters = argc > 2 ? atoi(argv[2]) : 30;
  int64_t* col;
  unsigned long long* out;
  cudaMalloc(&col, rows * 8);
  cudaMalloc(&out, 8);
  fill<<<148 * 8, 256>>>(col, rows);
  cudaDeviceSynchronize();
  run<0>(col, rows, out, iters); run<1>(col, rows, out, iters); run<2>(col, rows, out, iters);
  run<3>(col, rows, out, iters); run<4>(col, rows, out, iters); run<5>(col, rows, out, iters);
  return 0;
}
