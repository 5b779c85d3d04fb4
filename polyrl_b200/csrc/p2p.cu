// polyrl_b200 — NVLink/NVSwitch peer-memory kernels (symmetric heap = cudaMalloc + CUDA IPC, every
// rank holds device pointers to every peer's buffer).
//
//  * barrier:            device-side all-rank barrier on monotone epoch flags (st.release.sys /
//                        ld.acquire.sys) — replaces the reference's dist.barrier() fences around weight
//                        copies (rlboost/verl_stream/workers/stream_fsdp_workers.py:245,249,477,485)
//  * weight push:        trainer shard -> peer rollout replica HBM, optionally fused with row-wise fp8
//                        quantisation — replaces the whole shm + TCP sendfile + ZMQ + H2D + broadcast
//                        chain (rlboost/weight_transfer/*, rlboost/sglang/patches.py:137-357)
//  * all-gather pull:    each rank pulls every peer's shard (rollout weight sync / ZeRO-3 unshard)
//  * reduce-scatter pull fused with fp32 accumulate + grad-norm partial: the owner of a shard sums the
//                        peers' bf16 gradients straight out of their HBM into its fp32 accumulator —
//                        replaces NCCL reduce_scatter + cast/scale + norm
#include "common.cuh"

namespace prl {

constexpr int kMaxPeers = 16;
struct PeerPtrs {
  void* p[kMaxPeers];
};

// flags layout per rank: uint32 flags[kMaxPeers]; rank r writes `epoch` into flags[r] on every peer.
// The epoch lives in DEVICE memory (one counter per barrier channel, bumped by the kernel itself): a barrier captured into a
// CUDA graph therefore advances on every replay — a host-assigned epoch would be frozen at capture time and later replays
// would fall straight through (found on B200: TP engines lost lock-step under graph replay).
__global__ void p2p_barrier_kernel(PeerPtrs flag_ptrs, int rank, int world, uint32_t* __restrict__ epoch_ctr) {
  const int t = threadIdx.x;
  __shared__ uint32_t s_epoch;
  if (t == 0) { s_epoch = *epoch_ctr + 1; *epoch_ctr = s_epoch; }
  __syncthreads();
  const uint32_t epoch = s_epoch;
  if (t < world) {
    __threadfence_system();
    // max (not store): barriers issued from different streams of one process may execute out of launch order
    asm volatile("red.release.sys.global.max.u32 [%0], %1;" ::"l"(static_cast<uint32_t*>(flag_ptrs.p[t]) + rank), "r"(epoch) : "memory");
  }
  if (t < world) wait_ge_sys(static_cast<uint32_t*>(flag_ptrs.p[rank]) + t, epoch);
  __syncthreads();
}

__global__ void p2p_copy_kernel(const int4* __restrict__ src, int4* __restrict__ dst, int64_t n16) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x;
  for (; i + 3 * stride < n16; i += 4 * stride) {
    const int4 a = ld_v4(src + i), b = ld_v4(src + i + stride), c = ld_v4(src + i + 2 * stride),
               d = ld_v4(src + i + 3 * stride);
    st_v4(dst + i, a); st_v4(dst + i + stride, b); st_v4(dst + i + 2 * stride, c); st_v4(dst + i + 3 * stride, d);
  }
  for (; i < n16; i += stride) st_v4(dst + i, ld_v4(src + i));
}

// Row-wise e4m3 quantise a bf16 [rows, K] shard and store payload + scale into (peer) memory.
__global__ void __launch_bounds__(256)
p2p_push_fp8_kernel(const __nv_bfloat16* __restrict__ src, uint8_t* __restrict__ dst_q, float* __restrict__ dst_scale,
                    int rows, int K) {
  __shared__ float red[8];
  for (int row = blockIdx.x; row < rows; row += gridDim.x) {
    const __nv_bfloat16* xr = src + static_cast<int64_t>(row) * K;
    float amax = 0.f;
    for (int i = threadIdx.x * 8; i < K; i += 256 * 8) {
      const int4 v = ld_nc_v4(xr + i);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = unpack_bf16(u[j]);
        amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
      }
    }
    amax = warp_max(amax);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = amax;
    __syncthreads();
    amax = red[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) amax = fmaxf(amax, red[i]);
    const float sc = fmaxf(amax, 1e-12f) / 448.f;
    const float inv = 1.f / sc;
    if (threadIdx.x == 0) dst_scale[row] = sc;
    uint8_t* qr = dst_q + static_cast<int64_t>(row) * K;
    for (int i = threadIdx.x * 8; i < K; i += 256 * 8) {
      const int4 v = ld_nc_v4(xr + i);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&v);
      uint32_t o[2];
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const float2 a = unpack_bf16(u[2 * j]), b = unpack_bf16(u[2 * j + 1]);
        uint16_t lo, hi;
        asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(lo) : "f"(a.y * inv), "f"(a.x * inv));
        asm("cvt.rn.satfinite.e4m3x2.f32 %0, %1, %2;" : "=h"(hi) : "f"(b.y * inv), "f"(b.x * inv));
        o[j] = static_cast<uint32_t>(lo) | (static_cast<uint32_t>(hi) << 16);
      }
      *reinterpret_cast<uint2*>(qr + i) = make_uint2(o[0], o[1]);
    }
  }
}

// out[r * n16 + i] = peer_r[i]  for every rank r (local rank copies from itself too).
__global__ void p2p_allgather_pull_kernel(PeerPtrs shards, int4* __restrict__ out, int64_t n16, int rank, int world) {
  const int64_t total = n16 * world;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < total; i += stride) {
    const int64_t rr = i / n16, off = i - rr * n16;
    const int r = static_cast<int>((rr + rank) % world);   // start with the local shard, then ring order
    st_v4(out + r * n16 + off, ld_v4(static_cast<const int4*>(shards.p[r]) + off));
  }
}

// acc[i] += scale * sum_r grads_r[shard_off + i]   (bf16 peer grads, fp32 accumulate), optional sumsq.
// Bandwidth notes (profiles/: the first version kept ONE 16-byte load per peer in flight per thread and reached 49 GB/s of
// NVLink ingest): every thread now owns U = 4 consecutive-in-stride vectors, issues the U loads of two peers (and the fp32
// accumulator loads) before the first add, so >= 8 x 16 B are in flight per thread and ~1 MB per 32-CTA launch.
template <int U>
__device__ __forceinline__ void rs_accum_vec(const PeerPtrs& grads, int64_t shard_off, float* __restrict__ acc, const int64_t (&idx)[U],
                                             int nvalid, float scale, int world, int rank, float& local) {
  float s[U][8];
#pragma unroll
  for (int u = 0; u < U; ++u)
#pragma unroll
    for (int j = 0; j < 8; ++j) s[u][j] = 0.f;
  float4 a0[U], a1[U];
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u < nvalid) {
      a0[u] = *reinterpret_cast<const float4*>(acc + 8 * idx[u]);
      a1[u] = *reinterpret_cast<const float4*>(acc + 8 * idx[u] + 4);
    }
  for (int rr = 0; rr < world; rr += 2) {
    int4 v[2][U];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int r = (rr + h + rank) % world;
      const __nv_bfloat16* base = static_cast<const __nv_bfloat16*>(grads.p[r]) + shard_off;
#pragma unroll
      for (int u = 0; u < U; ++u)
        v[h][u] = (rr + h < world && u < nvalid) ? ld_v4(base + 8 * idx[u]) : make_int4(0, 0, 0, 0);
    }
#pragma unroll
    for (int h = 0; h < 2; ++h)
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t* w = reinterpret_cast<const uint32_t*>(&v[h][u]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16(w[j]);
          s[u][2 * j] += f.x; s[u][2 * j + 1] += f.y;
        }
      }
  }
#pragma unroll
  for (int u = 0; u < U; ++u)
    if (u < nvalid) {
      float4 x = a0[u], y = a1[u];
      x.x += scale * s[u][0]; x.y += scale * s[u][1]; x.z += scale * s[u][2]; x.w += scale * s[u][3];
      y.x += scale * s[u][4]; y.y += scale * s[u][5]; y.z += scale * s[u][6]; y.w += scale * s[u][7];
      *reinterpret_cast<float4*>(acc + 8 * idx[u]) = x;
      *reinterpret_cast<float4*>(acc + 8 * idx[u] + 4) = y;
      local += x.x * x.x + x.y * x.y + x.z * x.z + x.w * x.w + y.x * y.x + y.y * y.y + y.z * y.z + y.w * y.w;
    }
}

__global__ void __launch_bounds__(512)
p2p_reduce_scatter_accum_kernel(PeerPtrs grads, int64_t shard_off, float* __restrict__ acc, int64_t n, float scale, int world,
                                int rank, float* __restrict__ sumsq) {
  constexpr int U = 4;
  const int64_t n8 = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float local = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8; i += U * stride) {
    int64_t idx[U];
    int nvalid = 0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      idx[u] = i + u * stride;
      nvalid += idx[u] < n8;
    }
    rs_accum_vec<U>(grads, shard_off, acc, idx, nvalid, scale, world, rank, local);
  }
  if (sumsq) {
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) atomicAdd(sumsq, local);
  }
}

// One-shot all-reduce for small tensors (TP decode activations, <= a few MB): every rank has written its bf16 partial into its
// symmetric buffer (and a device barrier has passed); each rank reads ALL peers' partials over NVLink, sums in fp32 IN RANK
// ORDER (bit-identical results on every rank: TP engines stay in lock-step without host traffic), adds the residual and
// writes its own output.  out may alias residual.
__global__ void __launch_bounds__(256)
p2p_allreduce_kernel(PeerPtrs bufs, int64_t elem_off, __nv_bfloat16* __restrict__ out, const __nv_bfloat16* __restrict__ residual,
                     int64_t n, int world) {
  const int64_t n8 = n >> 3;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    int4 v[kMaxPeers];
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < world) v[r] = ld_v4(static_cast<const __nv_bfloat16*>(bufs.p[r]) + elem_off + 8 * i);
    float s[8];
    if (residual) {
      const int4 rv = ld_v4(residual + 8 * i);
      const uint32_t* u = reinterpret_cast<const uint32_t*>(&rv);
#pragma unroll
      for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(u[j]); s[2 * j] = f.x; s[2 * j + 1] = f.y; }
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = 0.f;
    }
#pragma unroll
    for (int r = 0; r < kMaxPeers; ++r)
      if (r < world) {
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&v[r]);
#pragma unroll
        for (int j = 0; j < 4; ++j) { const float2 f = unpack_bf16(u[j]); s[2 * j] += f.x; s[2 * j + 1] += f.y; }
      }
    int4 o;
    uint32_t* ou = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
    for (int j = 0; j < 4; ++j) ou[j] = pack_bf16(s[2 * j], s[2 * j + 1]);
    *reinterpret_cast<int4*>(out + 8 * i) = o;
  }
}

}  // namespace prl

using namespace prl;

extern "C" int prl_ipc_get_handle(void* dptr, void* out64) {
  cudaIpcMemHandle_t h;
  PRL_CUDA_CHECK(cudaIpcGetMemHandle(&h, dptr));
  memcpy(out64, &h, sizeof(h));
  return 0;
}
extern "C" int prl_ipc_open_handle(const void* in64, void** out_ptr) {
  cudaIpcMemHandle_t h;
  memcpy(&h, in64, sizeof(h));
  PRL_CUDA_CHECK(cudaIpcOpenMemHandle(out_ptr, h, cudaIpcMemLazyEnablePeerAccess));
  return 0;
}
extern "C" int prl_ipc_close_handle(void* ptr) {
  PRL_CUDA_CHECK(cudaIpcCloseMemHandle(ptr));
  return 0;
}
extern "C" int prl_cuda_malloc(void** out, int64_t bytes) {
  PRL_CUDA_CHECK(cudaMalloc(out, static_cast<size_t>(bytes)));
  PRL_CUDA_CHECK(cudaMemset(*out, 0, static_cast<size_t>(bytes)));
  return 0;
}
extern "C" int prl_cuda_free(void* p) {
  PRL_CUDA_CHECK(cudaFree(p));
  return 0;
}
extern "C" int prl_enable_peer_access(int peer_dev) {
  int can = 0, dev = 0;
  PRL_CUDA_CHECK(cudaGetDevice(&dev));
  if (peer_dev == dev) return 0;
  PRL_CUDA_CHECK(cudaDeviceCanAccessPeer(&can, dev, peer_dev));
  if (!can) return -40;
  cudaError_t e = cudaDeviceEnablePeerAccess(peer_dev, 0);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return 0; }
  return e == cudaSuccess ? 0 : static_cast<int>(e);
}

static PeerPtrs to_ptrs(void* const* ptrs, int world) {
  PeerPtrs pp;
  for (int i = 0; i < kMaxPeers; ++i) pp.p[i] = i < world ? ptrs[i] : nullptr;
  return pp;
}

extern "C" int prl_p2p_barrier(void* const* flag_ptrs, int rank, int world, uint32_t* epoch_ctr, cudaStream_t s) {
  if (world > kMaxPeers) return -41;
  p2p_barrier_kernel<<<1, 32, 0, s>>>(to_ptrs(flag_ptrs, world), rank, world, epoch_ctr);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int prl_p2p_copy(const void* src, void* dst, int64_t nbytes, int max_ctas, cudaStream_t s) {
  if (nbytes % 16 != 0) return -42;
  const int64_t n16 = nbytes / 16;
  if (n16 == 0) return 0;
  int grid = static_cast<int>((n16 + 1023) / 1024);
  const int cap = max_ctas > 0 ? max_ctas : 148 * 4;
  if (grid > cap) grid = cap;
  p2p_copy_kernel<<<grid, 256, 0, s>>>(static_cast<const int4*>(src), static_cast<int4*>(dst), n16);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int prl_p2p_push_fp8(const void* src, void* dst_q, float* dst_scale, int rows, int K, int max_ctas,
                                cudaStream_t s) {
  if (K % 8 != 0) return -43;
  if (rows <= 0) return 0;
  int grid = rows;
  const int cap = max_ctas > 0 ? max_ctas : 148 * 8;
  if (grid > cap) grid = cap;
  p2p_push_fp8_kernel<<<grid, 256, 0, s>>>(static_cast<const __nv_bfloat16*>(src), static_cast<uint8_t*>(dst_q),
                                           dst_scale, rows, K);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int prl_p2p_allgather_pull(void* const* shard_ptrs, void* out, int64_t shard_bytes, int rank, int world,
                                      int max_ctas, cudaStream_t s) {
  if (world > kMaxPeers || shard_bytes % 16 != 0) return -44;
  const int64_t n16 = shard_bytes / 16;
  if (n16 == 0) return 0;
  int grid = static_cast<int>((n16 * world + 1023) / 1024);
  const int cap = max_ctas > 0 ? max_ctas : 148 * 4;
  if (grid > cap) grid = cap;
  p2p_allgather_pull_kernel<<<grid, 256, 0, s>>>(to_ptrs(shard_ptrs, world), static_cast<int4*>(out), n16, rank, world);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int prl_p2p_allreduce(void* const* buf_ptrs, int64_t elem_off, void* out, const void* residual, int64_t n, int world,
                                 int max_ctas, cudaStream_t s) {
  if (world > kMaxPeers || n % 8 != 0) return -46;
  if (n == 0) return 0;
  int grid = static_cast<int>((n / 8 + 255) / 256);       // one 16-byte vector per thread: a single NVLink round trip deep
  const int cap = max_ctas > 0 ? max_ctas : 148 * 4;
  if (grid > cap) grid = cap;
  p2p_allreduce_kernel<<<grid, 256, 0, s>>>(to_ptrs(buf_ptrs, world), elem_off, static_cast<__nv_bfloat16*>(out),
                                            static_cast<const __nv_bfloat16*>(residual), n, world);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int prl_p2p_reduce_scatter_accum(void* const* grad_ptrs, int64_t shard_off, float* acc, int64_t n,
                                            float scale, int world, int rank, float* sumsq, int max_ctas,
                                            cudaStream_t s) {
  if (world > kMaxPeers || n % 8 != 0 || shard_off % 8 != 0) return -45;
  if (n == 0) return 0;
  int grid = static_cast<int>((n / 8 + 2047) / 2048);
  if (grid < 1) grid = 1;
  const int cap = max_ctas > 0 ? max_ctas : 148 * 2;
  if (grid > cap) grid = cap;
  p2p_reduce_scatter_accum_kernel<<<grid, 512, 0, s>>>(to_ptrs(grad_ptrs, world), shard_off, acc, n, scale, world, rank,
                                                      sumsq);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
