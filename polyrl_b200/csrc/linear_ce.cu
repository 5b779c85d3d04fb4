// polyrl_b200 — fused LM-head: tcgen05 GEMM hidden·W^T with an online-softmax epilogue that yields
// per-token log-prob and entropy WITHOUT materialising the [T, V] logits, plus the recompute
// backward pass that emits d_logits for one vocab slice at a time.
//
// Capability parity: 3rdparty/verl/verl/utils/kernel/kernels.py:157-379 (Triton fwd mainloop +
// epilogue) and :1124-1375 (d_logits split-N backward) / linear_cross_entropy.py:119.  Math per
// m = max z, S = sum exp(z-m), B = sum z·exp(z-m)/S, logprob = z_label - m - log S,
// entropy = log S + m - B, z = h·W^T / tau.
#include "gemm2_sm100.cuh"

namespace prl {

constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct CeFwdParams {
  const int64_t* labels;   // [T]
  float inv_tau;
  float* pmax;             // [T, S] running max (base-2 domain: z*log2e)
  float* psum;             // [T, S] sum 2^(z2 - max)
  float* pdot;             // [T, S] sum z * 2^(z2 - max)
  float* label_logit;      // [T]   z[label]
  int S;
};

struct EpiCeFwd {
  using Params = CeFwdParams;
  float m2, s, b;
  int64_t label;
  __device__ __forceinline__ void begin(const Params& p, int row, int M, const WorkInfo&) {
    m2 = -INFINITY; s = 0.f; b = 0.f;
    label = (row < M) ? p.labels[row] : -1;
  }
  template <int BN>
  __device__ __forceinline__ void tile(const Params& p, const AccReader& acc, int row, int M, int n0, int N) {
    const float scale2 = p.inv_tau * kLog2e;
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      acc.load(c, v);
      const int col = n0 + c;
      const int valid = N - col;
      if (valid <= 0) continue;
      float cmax = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = (i < valid) ? v[i] * scale2 : -INFINITY;   // base-2 scaled logits
        cmax = fmaxf(cmax, v[i]);
      }
      const float nm = fmaxf(m2, cmax);
      const float resc = exp2f(m2 - nm);                  // m2 = -inf -> 0
      float ls = 0.f, lb = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float e = exp2f(v[i] - nm);                 // masked cols: 2^-inf = 0
        ls += e;
        lb += (i < valid) ? v[i] * e : 0.f;
      }
      s = s * resc + ls;
      b = b * resc + lb;
      m2 = nm;
      const int64_t rel = label - col;
      if (rel >= 0 && rel < 32 && row < M) {
        float z = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) z = (i == rel) ? v[i] : z;
        p.label_logit[row] = z * kLn2;                    // back to natural units
      }
    }
  }
  __device__ __forceinline__ void end(const Params& p, int row, int M, const WorkInfo& wi) {
    if (row < M) {
      const int64_t o = static_cast<int64_t>(row) * p.S + wi.split;
      p.pmax[o] = m2;
      p.psum[o] = s;
      p.pdot[o] = b * kLn2;   // sum z_nat * 2^(..)
    }
  }
};

// Combine split statistics -> logprob, entropy, lse, E[z].
__global__ void ce_combine_kernel(const float* __restrict__ pmax, const float* __restrict__ psum,
                                  const float* __restrict__ pdot, const float* __restrict__ label_logit,
                                  const int64_t* __restrict__ labels, int T, int S, int V,
                                  float* __restrict__ logprob, float* __restrict__ entropy, float* __restrict__ lse,
                                  float* __restrict__ ez) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= T) return;
  float m = -INFINITY;
  for (int i = 0; i < S; ++i) m = fmaxf(m, pmax[static_cast<int64_t>(row) * S + i]);
  float s = 0.f, b = 0.f;
  for (int i = 0; i < S; ++i) {
    const float w = exp2f(pmax[static_cast<int64_t>(row) * S + i] - m);
    s += psum[static_cast<int64_t>(row) * S + i] * w;
    b += pdot[static_cast<int64_t>(row) * S + i] * w;
  }
  const float log_s = logf(s);
  const float m_nat = m * kLn2;
  const float l = m_nat + log_s;
  const float e = b / s;
  const int64_t lab = labels[row];
  const bool ok = lab >= 0 && lab < V;
  lse[row] = l;
  ez[row] = e;
  logprob[row] = ok ? label_logit[row] - l : 0.f;
  entropy[row] = l - e;
}

// d_logits for the vocab slice [v0, v0+N): dz = inv_tau * ( g_lp*(1[j==label] - p) - g_ent*p*(z - E[z]) )
struct CeBwdParams {
  const int64_t* labels;
  const float* lse;
  const float* ez;
  const float* g_lp;     // may be null
  const float* g_ent;    // may be null
  float inv_tau;
  int v0;
  __nv_bfloat16* DZ;     // [T, lddz]
  int64_t lddz;
};
struct EpiCeBwd {
  using Params = CeBwdParams;
  float lse2, ezv, glp, gent;
  int64_t label;
  __device__ __forceinline__ void begin(const Params& p, int row, int M, const WorkInfo&) {
    const bool ok = row < M;
    lse2 = ok ? p.lse[row] * kLog2e : 0.f;
    ezv = ok ? p.ez[row] : 0.f;
    glp = (ok && p.g_lp) ? p.g_lp[row] : 0.f;
    gent = (ok && p.g_ent) ? p.g_ent[row] : 0.f;
    label = ok ? p.labels[row] - p.v0 : -1;
  }
  __device__ __forceinline__ void end(const Params&, int, int, const WorkInfo&) {}
  template <int BN>
  __device__ __forceinline__ void tile(const Params& p, const AccReader& acc, int row, int M, int n0, int N) {
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      acc.load(c, v);
      const int col = n0 + c;
      const int valid = min(32, N - col);
      if (row >= M || valid <= 0) continue;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        const float z = v[i] * p.inv_tau;
        const float pr = exp2f(z * kLog2e - lse2);
        const float onehot = (static_cast<int64_t>(col + i) == label) ? 1.f : 0.f;
        v[i] = p.inv_tau * (glp * (onehot - pr) - gent * pr * (z - ezv));
      }
      __nv_bfloat16* dst = p.DZ + static_cast<int64_t>(row) * p.lddz + col;
      if (valid >= 32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          int4 o;
          o.x = pack_bf16(v[8 * i + 0], v[8 * i + 1]);
          o.y = pack_bf16(v[8 * i + 2], v[8 * i + 3]);
          o.z = pack_bf16(v[8 * i + 4], v[8 * i + 5]);
          o.w = pack_bf16(v[8 * i + 6], v[8 * i + 7]);
          st_v4(dst + 8 * i, o);
        }
      } else {
        for (int i = 0; i < valid; ++i) dst[i] = __float2bfloat16(v[i]);
      }
    }
  }
};

}  // namespace prl

using namespace prl;

extern "C" int prl_linear_ce_num_splits(int T, int V) {
  // T > 128 runs on the CTA-pair kernel: 256-row m-blocks, one work item per SM PAIR
  const int mb = T > kBlockM ? (T + 255) / 256 : 1;
  const int nb = (V + 255) / 256;
  const int workers = T > kBlockM ? sm_count() / 2 : sm_count();
  int s = (2 * workers + mb - 1) / mb;   // ~2 waves of work items
  if (s < 1) s = 1;
  if (s > nb) s = nb;
  const int per = (nb + s - 1) / s;
  return (nb + per - 1) / per;
}

// hidden:[T,H] bf16, W:[V,H] bf16, labels:[T] int64 -> logprob/entropy/lse/ez:[T] fp32.
// scratch: 3*T*S + T floats.
extern "C" int prl_linear_ce_fwd(const void* hidden, const void* W, const int64_t* labels, float inv_tau, int T, int V,
                                 int H, int64_t ldh, int64_t ldw, float* scratch, int S, float* logprob,
                                 float* entropy, float* lse, float* ez, cudaStream_t stream) {
  if (T <= 0) return 0;
  CeFwdParams p;
  p.labels = labels;
  p.inv_tau = inv_tau;
  p.pmax = scratch;
  p.psum = scratch + static_cast<int64_t>(T) * S;
  p.pdot = scratch + 2 * static_cast<int64_t>(T) * S;
  p.label_logit = scratch + 3 * static_cast<int64_t>(T) * S;
  p.S = S;
  using Cfg = GemmCfg<256, 2>;
  using Cfg2 = Gemm2Cfg<256, false, false, false>;
  int rc = T > kBlockM
               ? launch_gemm2<Cfg2, EpiCeFwd>(hidden, W, nullptr, T, V, H, ldh * 2, ldw * 2, 0, p, S, stream)
               : launch_gemm<Cfg, EpiCeFwd, NoGate, false>(hidden, W, T, V, H, ldh * 2, ldw * 2, p, NoGate::Params{}, S, 0,
                                                           stream);
  if (rc) return rc;
  ce_combine_kernel<<<(T + 127) / 128, 128, 0, stream>>>(p.pmax, p.psum, p.pdot, p.label_logit, labels, T, S, V,
                                                         logprob, entropy, lse, ez);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

// d_logits slice: DZ[T, Vc] for vocab rows [v0, v0+Vc) of W.
extern "C" int prl_linear_ce_dz(const void* hidden, const void* W_slice, const int64_t* labels, const float* lse,
                                const float* ez, const float* g_lp, const float* g_ent, float inv_tau, int T, int Vc,
                                int H, int v0, int64_t ldh, int64_t ldw, void* DZ, int64_t lddz, cudaStream_t stream) {
  CeBwdParams p{labels, lse, ez, g_lp, g_ent, inv_tau, v0, static_cast<__nv_bfloat16*>(DZ), lddz};
  using Cfg = GemmCfg<256, 2>;
  using Cfg2 = Gemm2Cfg<256, false, false, false>;
  if (T > kBlockM) return launch_gemm2<Cfg2, EpiCeBwd>(hidden, W_slice, nullptr, T, Vc, H, ldh * 2, ldw * 2, 0, p, 0, stream);
  return launch_gemm<Cfg, EpiCeBwd, NoGate, false>(hidden, W_slice, T, Vc, H, ldh * 2, ldw * 2, p, NoGate::Params{}, 0,
                                                   0, stream);
}
