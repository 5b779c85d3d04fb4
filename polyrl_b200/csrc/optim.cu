// polyrl_b200 — optimizer-side kernels for the ZeRO-partitioned trainer.
//
//  * grad_accum:  fp32 shard accumulator += scale * bf16/fp32 gradient (streaming PPO keeps sharded
//                 grads alive across update_actor_stream calls; reference semantics
//                 rlboost/verl_stream/workers/actor/stream_dp_actor.py:153-230)
//  * sumsq:       per-shard sum of squares -> device scalar (global grad-norm partial)
//  * adamw:       fused clip-scale + AdamW on the fp32 master shard, bf16 parameter write-back,
//                 gradient zeroing, non-finite-norm => skip (reference:
//                 3rdparty/verl/verl/workers/actor/dp_actor.py:275-294 + torch.optim.AdamW)
// All scalars that depend on device results (total grad norm) are read from device memory so the
// optimizer step needs no host synchronisation and can sit inside a CUDA graph.
#include "common.cuh"

namespace prl {

template <typename G>
__device__ __forceinline__ float4 load4(const G* p);
template <>
__device__ __forceinline__ float4 load4<float>(const float* p) {
  return *reinterpret_cast<const float4*>(p);
}
template <>
__device__ __forceinline__ float4 load4<__nv_bfloat16>(const __nv_bfloat16* p) {
  const uint2 u = *reinterpret_cast<const uint2*>(p);
  const float2 a = unpack_bf16(u.x), b = unpack_bf16(u.y);
  return make_float4(a.x, a.y, b.x, b.y);
}

template <typename G>
__global__ void grad_accum_kernel(float* __restrict__ acc, const G* __restrict__ g, float scale, int64_t n,
                                  float* __restrict__ sumsq) {
  float local = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    float4 a = *reinterpret_cast<float4*>(acc + 4 * i);
    const float4 v = load4<G>(g + 4 * i);
    a.x += scale * v.x; a.y += scale * v.y; a.z += scale * v.z; a.w += scale * v.w;
    *reinterpret_cast<float4*>(acc + 4 * i) = a;
    local += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = n4 << 2; i < n; ++i) {
      float a = acc[i] + scale * static_cast<float>(g[i]);
      acc[i] = a;
      local += a * a;
    }
  }
  if (sumsq) {
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) atomicAdd(sumsq, local);
  }
}

__global__ void sumsq_kernel(const float* __restrict__ g, int64_t n, float* __restrict__ out) {
  float local = 0.f;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(g + 4 * i);
    local += a.x * a.x + a.y * a.y + a.z * a.z + a.w * a.w;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)
    for (int64_t i = n4 << 2; i < n; ++i) local += g[i] * g[i];
  local = warp_sum(local);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, local);
}

struct AdamArgs {
  float lr, beta1, beta2, eps, weight_decay, bc1, bc2;   // bc = 1 - beta^t
  float max_norm;                                         // <= 0 disables clipping
};

// total_sumsq: device scalar holding the GLOBAL sum of squared grads (already all-reduced).
// status[0] is set to 1 when the step was skipped because the norm is non-finite.
__global__ void adamw_kernel(float* __restrict__ master, float* __restrict__ m, float* __restrict__ v,
                             float* __restrict__ grad, __nv_bfloat16* __restrict__ param_bf16, int64_t n,
                             AdamArgs a, const float* __restrict__ total_sumsq, int* __restrict__ status,
                             int zero_grad) {
  float clip = 1.f;
  bool skip = false;
  if (total_sumsq) {
    const float norm = sqrtf(*total_sumsq);
    if (!isfinite(norm)) skip = true;
    else if (a.max_norm > 0.f) clip = fminf(1.f, a.max_norm / (norm + 1e-6f));
  }
  if (skip && blockIdx.x == 0 && threadIdx.x == 0 && status) status[0] = 1;
  const int64_t n4 = n >> 2;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n4;
       i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
    if (!skip) {
      float4 g = *reinterpret_cast<float4*>(grad + 4 * i);
      float4 p = *reinterpret_cast<float4*>(master + 4 * i);
      float4 mm = *reinterpret_cast<float4*>(m + 4 * i);
      float4 vv = *reinterpret_cast<float4*>(v + 4 * i);
      float* gp = reinterpret_cast<float*>(&g);
      float* pp = reinterpret_cast<float*>(&p);
      float* mp = reinterpret_cast<float*>(&mm);
      float* vp = reinterpret_cast<float*>(&vv);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float gj = gp[j] * clip;
        pp[j] *= (1.f - a.lr * a.weight_decay);
        mp[j] = a.beta1 * mp[j] + (1.f - a.beta1) * gj;
        vp[j] = a.beta2 * vp[j] + (1.f - a.beta2) * gj * gj;
        const float denom = sqrtf(vp[j] / a.bc2) + a.eps;
        pp[j] -= (a.lr / a.bc1) * mp[j] / denom;
      }
      *reinterpret_cast<float4*>(master + 4 * i) = p;
      *reinterpret_cast<float4*>(m + 4 * i) = mm;
      *reinterpret_cast<float4*>(v + 4 * i) = vv;
      if (param_bf16)
        *reinterpret_cast<uint2*>(param_bf16 + 4 * i) = make_uint2(pack_bf16(pp[0], pp[1]), pack_bf16(pp[2], pp[3]));
    }
    if (zero_grad) *reinterpret_cast<float4*>(grad + 4 * i) = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    for (int64_t i = n4 << 2; i < n; ++i) {
      if (!skip) {
        const float gj = grad[i] * clip;
        float p = master[i] * (1.f - a.lr * a.weight_decay);
        const float mj = a.beta1 * m[i] + (1.f - a.beta1) * gj;
        const float vj = a.beta2 * v[i] + (1.f - a.beta2) * gj * gj;
        p -= (a.lr / a.bc1) * mj / (sqrtf(vj / a.bc2) + a.eps);
        master[i] = p; m[i] = mj; v[i] = vj;
        if (param_bf16) param_bf16[i] = __float2bfloat16(p);
      }
      if (zero_grad) grad[i] = 0.f;
    }
  }
}

}  // namespace prl

using namespace prl;

static int grid_for(int64_t n4) {
  int64_t g = (n4 + 255) / 256;
  if (g > 148 * 8) g = 148 * 8;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

// g_dtype: 0 = fp32, 1 = bf16
extern "C" int prl_grad_accum(float* acc, const void* g, int g_dtype, float scale, int64_t n, float* sumsq,
                              cudaStream_t s) {
  if (n <= 0) return 0;
  if (g_dtype == 0)
    grad_accum_kernel<float><<<grid_for(n >> 2), 256, 0, s>>>(acc, static_cast<const float*>(g), scale, n, sumsq);
  else
    grad_accum_kernel<__nv_bfloat16><<<grid_for(n >> 2), 256, 0, s>>>(acc, static_cast<const __nv_bfloat16*>(g), scale, n, sumsq);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int prl_sumsq(const float* g, int64_t n, float* out, cudaStream_t s) {
  if (n <= 0) return 0;
  sumsq_kernel<<<grid_for(n >> 2), 256, 0, s>>>(g, n, out);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
extern "C" int prl_adamw(float* master, float* m, float* v, float* grad, void* param_bf16, int64_t n, float lr,
                         float beta1, float beta2, float eps, float weight_decay, int step, float max_norm,
                         const float* total_sumsq, int* status, int zero_grad, cudaStream_t s) {
  if (n <= 0) return 0;
  AdamArgs a;
  a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
  a.bc1 = 1.f - powf(beta1, static_cast<float>(step));
  a.bc2 = 1.f - powf(beta2, static_cast<float>(step));
  a.max_norm = max_norm;
  adamw_kernel<<<grid_for(n >> 2), 256, 0, s>>>(master, m, v, grad, static_cast<__nv_bfloat16*>(param_bf16), n, a,
                                                total_sumsq, status, zero_grad);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
