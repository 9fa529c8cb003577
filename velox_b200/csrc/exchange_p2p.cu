// Peer-memory exchange kernels: the data path of B200PartitionedOutput -> B200Exchange between the
// GPUs of one node without NCCL in the loop. Every rank owns an exchange heap (cudaMalloc) that all
// peers map with CUDA IPC; NVLink / NVSwitch carries plain stores into the destination's heap:
//
//   put_block   one small block (per-destination row counts + dictionaries) into every peer
//   put_rows    partition gather FUSED with the transfer: row j of the partition-grouped order goes
//               straight from the source columns into its destination's segment, column-major, so the
//               shuffle never materialises a send buffer and the stores of one destination are contiguous
//   signal      release-store of the epoch into every peer's flag word (after the data is fenced)
//   wait        acquire-spin until every source's flag reached the epoch (bounded: sets an error flag)
//   collect     the W source segments of this rank's heap -> contiguous output columns
//
// Replaces the serialise -> OutputBuffer -> HTTP/UCX pull of the reference
// (velox/exec/PartitionedOutput.cpp, velox/exec/ExchangeClient.cpp,
// velox/experimental/ucx-exchange/UcxPartitionedOutput.h) for ranks that share an NVLink domain.
#include "common.cuh"

namespace vb2 {

constexpr int kP2pMaxWorld = 16;
constexpr int kP2pMaxCols = 24;

struct P2pPeers {
  void* p[kP2pMaxWorld];
};
struct P2pCols {
  const void* src[kP2pMaxCols];
  int32_t width[kP2pMaxCols];
  int n;
};

__host__ __device__ __forceinline__ int64_t p2p_align(int64_t v) { return (v + 127) / 128 * 128; }
// byte offset of column c inside a segment that holds `count` rows (both sides compute it)
__host__ __device__ __forceinline__ int64_t p2p_col_offset(const int32_t* widths, int c, int64_t count) {
  int64_t off = 0;
  for (int k = 0; k < c; ++k) off += p2p_align(count * widths[k]);
  return off;
}

__device__ __forceinline__ void st_release_sys(uint64_t* p, uint64_t v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ uint64_t ld_acquire_sys(const uint64_t* p) {
  uint64_t v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ uint64_t global_timer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// block p copies `bytes` (multiple of 16) of src into peer p's destination
__global__ void p2p_put_block_kernel(const __grid_constant__ P2pPeers dst, const uint4* __restrict__ src, int64_t bytes) {
  uint4* d = reinterpret_cast<uint4*>(dst.p[blockIdx.x]);
  const int64_t n16 = bytes >> 4;
  for (int64_t i = threadIdx.x; i < n16; i += blockDim.x) d[i] = src[i];
  __threadfence_system();
}

__global__ void p2p_signal_kernel(const __grid_constant__ P2pPeers flags, int world, uint64_t epoch) {
  if (threadIdx.x < world) {
    __threadfence_system();
    st_release_sys(reinterpret_cast<uint64_t*>(flags.p[threadIdx.x]), epoch);
  }
}

__global__ void p2p_wait_kernel(const uint64_t* __restrict__ flags, int stride_words, int world, uint64_t epoch, int32_t* __restrict__ error_flag,
                                uint64_t timeout_ns) {
  if (threadIdx.x < world) {
    const uint64_t* f = flags + static_cast<int64_t>(threadIdx.x) * stride_words;
    const uint64_t t0 = global_timer_ns();
    while (ld_acquire_sys(f) < epoch) {
      __nanosleep(200);
      if (global_timer_ns() - t0 > timeout_ns) {
        atomicCAS(error_flag, 0, 200 + threadIdx.x);  // peer threadIdx.x never arrived
        break;
      }
    }
  }
  __syncthreads();
  __threadfence_system();
}

// Rows grouped by destination (order[j] = source row of the j-th grouped row, counts[p] rows for
// destination p) -> destination segments, all columns, one launch. broadcast: every destination
// receives all n rows in their original order.
__global__ void p2p_put_rows_kernel(const int32_t* __restrict__ order, const int64_t* __restrict__ counts, int world, int64_t n,
                                    const __grid_constant__ P2pCols cols, const __grid_constant__ P2pPeers seg, int broadcast) {
  __shared__ int64_t start[kP2pMaxWorld + 1];
  __shared__ int64_t col_off[kP2pMaxWorld][kP2pMaxCols];
  if (threadIdx.x == 0) {
    int64_t run = 0;
    for (int p = 0; p < world; ++p) {
      start[p] = run;
      run += broadcast ? 0 : counts[p];
    }
    start[world] = run;
  }
  for (int i = threadIdx.x; i < world * cols.n; i += blockDim.x) {
    const int p = i / cols.n, c = i % cols.n;
    col_off[p][c] = p2p_col_offset(cols.width, c, broadcast ? n : counts[p]);
  }
  __syncthreads();
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < n; j += stride) {
    const int64_t row = order ? order[j] : j;
    int p0 = 0, p1 = world;
    if (!broadcast) {
      int p = 0;
      while (p + 1 < world && j >= start[p + 1]) ++p;  // world <= 16: a short scan
      p0 = p;
      p1 = p + 1;
    }
    for (int c = 0; c < cols.n; ++c) {
      const int w = cols.width[c];
      uint64_t v;
      if (w == 8) v = reinterpret_cast<const uint64_t*>(cols.src[c])[row];
      else if (w == 4) v = reinterpret_cast<const uint32_t*>(cols.src[c])[row];
      else v = reinterpret_cast<const uint8_t*>(cols.src[c])[row];
      for (int p = p0; p < p1; ++p) {
        const int64_t pos = broadcast ? j : j - start[p];
        char* base = reinterpret_cast<char*>(seg.p[p]) + col_off[p][c];
        if (w == 8) reinterpret_cast<uint64_t*>(base)[pos] = v;
        else if (w == 4) reinterpret_cast<uint32_t*>(base)[pos] = static_cast<uint32_t>(v);
        else reinterpret_cast<uint8_t*>(base)[pos] = static_cast<uint8_t>(v);
      }
    }
  }
  __threadfence_system();
}

struct P2pCollect {
  const void* seg[kP2pMaxWorld];   // this rank's segment of every source
  int64_t count[kP2pMaxWorld];     // rows from every source
  int64_t row_start[kP2pMaxWorld];
  void* out[kP2pMaxCols];
  int32_t width[kP2pMaxCols];
  int world, ncols;
};
// blockIdx.y = source, blockIdx.z = column
__global__ void p2p_collect_kernel(const __grid_constant__ P2pCollect a) {
  const int s = blockIdx.y, c = blockIdx.z;
  const int64_t cnt = a.count[s];
  const int w = a.width[c];
  const char* src = reinterpret_cast<const char*>(a.seg[s]) + p2p_col_offset(a.width, c, cnt);
  char* dst = reinterpret_cast<char*>(a.out[c]) + a.row_start[s] * w;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < cnt; i += stride) {
    if (w == 8) reinterpret_cast<uint64_t*>(dst)[i] = reinterpret_cast<const uint64_t*>(src)[i];
    else if (w == 4) reinterpret_cast<uint32_t*>(dst)[i] = reinterpret_cast<const uint32_t*>(src)[i];
    else reinterpret_cast<uint8_t*>(dst)[i] = reinterpret_cast<const uint8_t*>(src)[i];
  }
}

}  // namespace vb2

using namespace vb2;

extern "C" {

int64_t vb2k_p2p_segment_bytes(const int32_t* widths, int32_t ncols, int64_t rows) { return p2p_col_offset(widths, ncols, rows); }

int vb2k_p2p_put_block(void* const* peer_dst, int32_t world, const void* src, int64_t bytes, void* stream) {
  if (world < 1 || world > kP2pMaxWorld || (bytes & 15)) return fail_msg(VB2_ERR_INVALID, "p2p_put_block: bad arguments");
  P2pPeers d{};
  for (int p = 0; p < world; ++p) d.p[p] = peer_dst[p];
  p2p_put_block_kernel<<<vb2::counted(world), 256, 0, static_cast<cudaStream_t>(stream)>>>(d, reinterpret_cast<const uint4*>(src), bytes);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_p2p_signal(void* const* peer_flags, int32_t world, uint64_t epoch, void* stream) {
  if (world < 1 || world > kP2pMaxWorld) return fail_msg(VB2_ERR_INVALID, "p2p_signal: bad world");
  P2pPeers f{};
  for (int p = 0; p < world; ++p) f.p[p] = peer_flags[p];
  p2p_signal_kernel<<<vb2::counted(1), 32, 0, static_cast<cudaStream_t>(stream)>>>(f, world, epoch);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_p2p_wait(const uint64_t* flags, int32_t stride_words, int32_t world, uint64_t epoch, int32_t* error_flag, uint64_t timeout_ns, void* stream) {
  if (world < 1 || world > kP2pMaxWorld) return fail_msg(VB2_ERR_INVALID, "p2p_wait: bad world");
  p2p_wait_kernel<<<vb2::counted(1), 32, 0, static_cast<cudaStream_t>(stream)>>>(flags, stride_words, world, epoch, error_flag, timeout_ns);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_p2p_put_rows(const int32_t* order, const int64_t* counts_dev, int32_t world, int64_t n, const void* const* cols, const int32_t* widths,
                      int32_t ncols, void* const* peer_segments, int32_t broadcast, void* stream) {
  if (world < 1 || world > kP2pMaxWorld || ncols < 1 || ncols > kP2pMaxCols) return fail_msg(VB2_ERR_INVALID, "p2p_put_rows: bad arguments");
  if (n <= 0) return VB2_OK;
  P2pCols c{};
  c.n = ncols;
  for (int i = 0; i < ncols; ++i) {
    if (widths[i] != 1 && widths[i] != 4 && widths[i] != 8) return fail_msg(VB2_ERR_INVALID, "p2p_put_rows: widths 1, 4, 8");
    c.src[i] = cols[i];
    c.width[i] = widths[i];
  }
  P2pPeers s{};
  for (int p = 0; p < world; ++p) s.p[p] = peer_segments[p];
  int64_t b = (n + 255) / 256, cap = static_cast<int64_t>(device_sm_count()) * 8;
  p2p_put_rows_kernel<<<vb2::counted(static_cast<unsigned>(b > cap ? cap : b)), 256, 0, static_cast<cudaStream_t>(stream)>>>(order, counts_dev, world, n, c, s,
                                                                                                                            broadcast);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

int vb2k_p2p_collect(const void* const* local_segments, const int64_t* counts, int32_t world, const int32_t* widths, int32_t ncols, void* const* outs,
                     void* stream) {
  if (world < 1 || world > kP2pMaxWorld || ncols < 1 || ncols > kP2pMaxCols) return fail_msg(VB2_ERR_INVALID, "p2p_collect: bad arguments");
  P2pCollect a{};
  a.world = world;
  a.ncols = ncols;
  int64_t run = 0, most = 0;
  for (int s = 0; s < world; ++s) {
    a.seg[s] = local_segments[s];
    a.count[s] = counts[s];
    a.row_start[s] = run;
    run += counts[s];
    most = counts[s] > most ? counts[s] : most;
  }
  for (int c = 0; c < ncols; ++c) { a.out[c] = outs[c]; a.width[c] = widths[c]; }
  if (most <= 0) return VB2_OK;
  int64_t bx = (most + 255) / 256, cap = static_cast<int64_t>(device_sm_count()) * 4;
  dim3 grid(static_cast<unsigned>(bx > cap ? cap : bx), static_cast<unsigned>(world), static_cast<unsigned>(ncols));
  vb2::note_launch();
  p2p_collect_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(a);
  VB2_CUDA_OK(cudaGetLastError());
  return VB2_OK;
}

}  // extern "C"
