// polyrl_b200 — token sampling for the rollout engine: temperature + Gumbel-max categorical sampling
// (exact for top_k=-1/top_p=1, the GRPO default) or greedy, returning the token id AND its
// log-probability under the temperature-scaled softmax in a single pass over the logits row.
//
// Reference capability: SGLang sampler + `return_logprob` (meta_info.output_token_logprobs consumed at
// rlboost/verl_stream/workers/rollout/sglang_rollout/sglang_rollout_remote.py:318-391).
#include "common.cuh"

namespace prl {

__device__ __forceinline__ uint32_t hash_u32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ float gumbel_noise(uint64_t seed, uint32_t row, uint32_t col) {
  uint32_t h = hash_u32(static_cast<uint32_t>(seed) ^ hash_u32(col + 0x9e3779b9U * (row + 1)));
  h = hash_u32(h ^ static_cast<uint32_t>(seed >> 32));
  const float u = (static_cast<float>(h >> 8) + 0.5f) * (1.0f / 16777216.0f);   // (0,1)
  return -__logf(-__logf(u));
}

struct SampleState {
  float best;      // max of z + gumbel
  int best_idx;
  float best_z;
  float m;         // running max of z
  float s;         // running sum exp(z - m)
};
__device__ __forceinline__ void merge(SampleState& a, const SampleState& b) {
  if (b.best > a.best || (b.best == a.best && b.best_idx < a.best_idx)) {
    a.best = b.best; a.best_idx = b.best_idx; a.best_z = b.best_z;
  }
  const float nm = fmaxf(a.m, b.m);
  if (nm == -INFINITY) return;
  a.s = a.s * __expf(a.m - nm) + b.s * __expf(b.m - nm);
  a.m = nm;
}

template <typename T>
__global__ void __launch_bounds__(1024)
sample_kernel(const T* __restrict__ logits, int64_t ld, int V, const float* __restrict__ inv_temp,
              const uint64_t* __restrict__ seeds, int greedy_all, int64_t* __restrict__ out_tok,
              float* __restrict__ out_lp) {
  const int row = blockIdx.x;
  const T* lr = logits + static_cast<int64_t>(row) * ld;
  const float it = inv_temp ? inv_temp[row] : 1.f;
  const bool greedy = greedy_all || it <= 0.f;
  const float scale = greedy ? 1.f : it;
  const uint64_t seed = seeds ? seeds[row] : 0;
  SampleState st{-INFINITY, 0x7fffffff, 0.f, -INFINITY, 0.f};
  auto visit = [&](int c, float raw) {
    const float z = raw * scale;
    const float key = greedy ? z : z + gumbel_noise(seed, row, c);
    if (key > st.best) { st.best = key; st.best_idx = c; st.best_z = z; }
    const float nm = fmaxf(st.m, z);
    st.s = st.s * __expf(st.m - nm) + __expf(z - nm);
    st.m = nm;
  };
  int c0 = 0;
  if constexpr (sizeof(T) == 2) {
    // 128-bit vector loads (8 bf16) when the row is 16-byte aligned: 8x fewer load instructions (profile: the scalar
    // version took 131 us per step for [128, 128256] logits)
    if ((reinterpret_cast<uintptr_t>(lr) & 15) == 0) {
      const int nvec = V >> 3;
      for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        const int4 q = ld_nc_v4(reinterpret_cast<const __nv_bfloat16*>(lr) + 8 * v);
        const uint32_t* u = reinterpret_cast<const uint32_t*>(&q);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = unpack_bf16(u[j]);
          visit(8 * v + 2 * j, f.x);
          visit(8 * v + 2 * j + 1, f.y);
        }
      }
      c0 = nvec << 3;
    }
  }
  for (int c = c0 + threadIdx.x; c < V; c += blockDim.x) visit(c, static_cast<float>(lr[c]));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    SampleState other;
    other.best = __shfl_xor_sync(0xffffffffu, st.best, o);
    other.best_idx = __shfl_xor_sync(0xffffffffu, st.best_idx, o);
    other.best_z = __shfl_xor_sync(0xffffffffu, st.best_z, o);
    other.m = __shfl_xor_sync(0xffffffffu, st.m, o);
    other.s = __shfl_xor_sync(0xffffffffu, st.s, o);
    merge(st, other);
  }
  __shared__ SampleState red[32];
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 0) red[w] = st;
  __syncthreads();
  if (threadIdx.x == 0) {
    SampleState t = red[0];
    for (int i = 1; i < (blockDim.x >> 5); ++i) merge(t, red[i]);
    out_tok[row] = t.best_idx;
    out_lp[row] = t.best_z - (t.m + __logf(t.s));
  }
}

}  // namespace prl

using namespace prl;

// dtype: 0 = fp32 logits, 1 = bf16 logits
extern "C" int prl_sample(const void* logits, int dtype, int64_t ld, int B, int V, const float* inv_temp,
                          const uint64_t* seeds, int greedy_all, int64_t* out_tok, float* out_lp, cudaStream_t s) {
  if (B <= 0) return 0;
  if (dtype == 0)
    sample_kernel<float><<<B, 1024, 0, s>>>(static_cast<const float*>(logits), ld, V, inv_temp, seeds, greedy_all,
                                           out_tok, out_lp);
  else
    sample_kernel<__nv_bfloat16><<<B, 1024, 0, s>>>(static_cast<const __nv_bfloat16*>(logits), ld, V, inv_temp, seeds,
                                                   greedy_all, out_tok, out_lp);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
