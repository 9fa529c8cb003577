// Device-side helpers shared by every kernel of the B200 operator path.
//
// Reference semantics restated here (file:line into /root/reference/velox):
//   NaN-aware comparisons ... type/FloatingPointUtil.h:52-98, functions/prestosql/Comparisons.h:24-160
//   checked integer math .... common/base/CheckedArithmetic.h:27-60
//   key hashing ............. exec/VectorHasher.cpp:62-126 (folly::hasher<T>), common/base/BitUtil.h:775-784
//                             (hashMix), common/base/BitUtil.cpp:177-230 (hashBytes, CRC32-C)
#pragma once
#ifdef __CUDACC_RTC__
// NVRTC (fused_jit.cu compiles fused_scan.cuh at run time): no system headers; the few names the
// device code needs from them are spelled out, and checked against the host's values by fused_jit.cu.
typedef long long int64_t;
typedef unsigned long long uint64_t;
typedef int int32_t;
typedef unsigned int uint32_t;
typedef unsigned char uint8_t;
typedef unsigned long long uintptr_t;
typedef unsigned long size_t;
#define INT64_MIN (-9223372036854775807LL - 1)
#define INT64_MAX 9223372036854775807LL
#define INT32_MIN (-2147483647 - 1)
#define INT32_MAX 2147483647
#define VB2_FUSED_MAX_COLS 8
#define VB2_FUSED_MAX_PARAMS 12
#define VB2_FUSED_MAX_KEYS 2
#else
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/velox_b200_kernels.h"
#endif

namespace vb2 {

constexpr int kWarp = 32;
constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

#ifndef __CUDACC_RTC__
#define VB2_CUDA_OK(expr)                                   \
  do {                                                      \
    cudaError_t _e = (expr);                                \
    if (_e != cudaSuccess) return vb2::fail_cuda(_e, #expr); \
  } while (0)

int fail_cuda(cudaError_t e, const char* what);  // records message, returns VB2_ERR_CUDA
int fail_msg(int code, const char* msg);
int device_sm_count();
// Every kernel launch of this library goes through counted(): vb2k_kernel_launches() is the claim
// bench.py reports as gpu_launches.
void note_launch();
template <class T>
inline T counted(T grid) {
  note_launch();
  return grid;
}
#endif

// ---------------------------------------------------------------------------------------------
// Streaming loads. Input columns are read exactly once: bypass L1 allocation, 128-bit wide.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double2 ldg_stream_f64x2(const double* p) {
  double2 r;
  asm("ld.global.nc.L1::no_allocate.L2::128B.v2.f64 {%0, %1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ int4 ldg_stream_i32x4(const int32_t* p) {
  int4 r;
  asm("ld.global.nc.L1::no_allocate.L2::128B.v4.s32 {%0, %1, %2, %3}, [%4];"
      : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
      : "l"(p));
  return r;
}
__device__ __forceinline__ longlong2 ldg_stream_i64x2(const int64_t* p) {
  longlong2 r;
  asm("ld.global.nc.L1::no_allocate.L2::128B.v2.s64 {%0, %1}, [%2];" : "=l"(r.x), "=l"(r.y) : "l"(p));
  return r;
}

// Row-level scalar semantics shared by the kernels, the expression interpreter and the expression
// JIT (which compiles the same text with NVRTC): validity bits, NaN-aware comparisons, checked
// integer arithmetic, LIKE and string comparison.
#include "vm_ops.inc"

// ---------------------------------------------------------------------------------------------
// Hashing (bit-exact with folly::hasher<T> as used by VectorHasher).
// ---------------------------------------------------------------------------------------------
constexpr uint64_t kNullHash = 1;

__host__ __device__ __forceinline__ uint64_t twang_mix64(uint64_t key) {
  key = (~key) + (key << 21);
  key = key ^ (key >> 24);
  key = key + (key << 3) + (key << 8);
  key = key ^ (key >> 14);
  key = key + (key << 2) + (key << 4);
  key = key ^ (key >> 28);
  key = key + (key << 31);
  return key;
}
__host__ __device__ __forceinline__ uint32_t jenkins_rev_mix32(uint32_t key) {
  key += (key << 12);
  key ^= (key >> 22);
  key += (key << 4);
  key ^= (key >> 9);
  key += (key << 10);
  key ^= (key >> 2);
  key += (key << 7);
  key += (key << 12);
  return key;
}
__host__ __device__ __forceinline__ uint64_t hash_mix(uint64_t upper, uint64_t lower) {
  const uint64_t kMul = 0x9ddfea08eb382d69ULL;
  uint64_t a = (lower ^ upper) * kMul;
  a ^= (a >> 47);
  uint64_t b = (upper ^ a) * kMul;
  b ^= (b >> 47);
  b *= kMul;
  return b;
}
__device__ __forceinline__ uint64_t hash_f64(double v) {
  if (isnan(v)) v = __longlong_as_double(0x7ff8000000000000LL);  // std::numeric_limits<double>::quiet_NaN()
  if (v == 0.0) return 0;                                        // +0 / -0 hash alike
  return twang_mix64(static_cast<uint64_t>(__double_as_longlong(v)));
}

// CRC32-C (Castagnoli, reflected polynomial 0x82F63B78), the function SSE4.2 crc32 implements.
// Bitwise form: the string-key path is latency- not throughput-bound on device.
__device__ __forceinline__ uint32_t crc32c_u8(uint32_t crc, uint8_t b) {
  crc ^= b;
#pragma unroll
  for (int k = 0; k < 8; ++k) crc = (crc >> 1) ^ (0x82F63B78u & (0u - (crc & 1u)));
  return crc;
}
// _mm_crc32_u64(seed, v): only the low 32 bits of seed take part; result zero-extended.
__device__ __forceinline__ uint64_t crc32c_u64(uint64_t seed, uint64_t v) {
  uint32_t crc = static_cast<uint32_t>(seed);
#pragma unroll
  for (int i = 0; i < 8; ++i) crc = crc32c_u8(crc, static_cast<uint8_t>(v >> (8 * i)));
  return crc;
}
__device__ __forceinline__ uint64_t load_partial_word(const uint8_t* p, int n) {
  uint64_t r = 0;
  for (int i = 0; i < n; ++i) r |= static_cast<uint64_t>(p[i]) << (8 * i);
  return r;
}
__device__ __forceinline__ uint64_t load_word(const uint8_t* p) { return load_partial_word(p, 8); }

__device__ inline uint64_t hash_bytes(uint64_t seed, const uint8_t* data, int32_t size) {
  const uint64_t kMul = 0x9ddfea08eb382d69ULL;
  if (size < 8) {
    uint64_t word = load_partial_word(data, size);
    uint64_t crc = crc32c_u64(seed, word);
    uint64_t crc2 = crc32c_u64(seed, word >> 32);
    return crc | (crc2 << 32);
  }
  uint64_t a0 = seed, a1 = seed << 32, a2 = seed >> 16;
  int32_t toGo = size;
  const uint8_t* p = data;
  while (toGo >= 24) {
    a0 = crc32c_u64(a0, load_word(p));
    a1 = crc32c_u64(a1, load_word(p + 8));
    a2 = crc32c_u64(a2, load_word(p + 16));
    p += 24;
    toGo -= 24;
  }
  if (toGo > 16) {
    a0 = crc32c_u64(a0, load_word(p));
    a1 = crc32c_u64(a1, load_word(p + 8));
    a2 = crc32c_u64(a2, load_partial_word(p + 16, toGo - 16));
  } else if (toGo > 8) {
    a0 = crc32c_u64(a0, load_word(p));
    a1 = crc32c_u64(a1, toGo == 16 ? load_word(p + 8) : load_partial_word(p + 8, toGo - 8));
  } else if (toGo > 0) {
    a0 = crc32c_u64(a0, toGo == 8 ? load_word(p) : load_partial_word(p, toGo));
  }
  return a0 ^ (a1 * kMul) ^ (a2 * kMul);
}

// ---------------------------------------------------------------------------------------------
// Warp / block reductions
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int64_t warp_sum(int64_t v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ int warp_sum(int v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace vb2
