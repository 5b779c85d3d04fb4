// polyrl_b200 — entry points of the CTA-pair (cta_group::2) tcgen05 GEMM family (gemm2_sm100.cuh):
// plain / residual GEMM in the NT, NN and TN layouts, i.e. the forward, dgrad and wgrad of a Linear
// layer, all on one mainloop with a TMA-store epilogue.  Replaces the cuBLAS calls the reference
// reaches through nn.Linear (3rdparty/verl/verl/workers/actor/dp_actor.py:86-273).
#include "gemm2_sm100.cuh"

namespace prl {

template <int BN, bool A_MN, bool B_MN>
static int run2(const void* A, const void* B, void* D, const void* R, int M, int N, int K, int64_t lda, int64_t ldb,
                int64_t ldd, cudaStream_t s) {
  using Cfg = Gemm2Cfg<BN, A_MN, B_MN, true>;
  TmaStoreParams ep{static_cast<const __nv_bfloat16*>(R), ldd};
  return launch_gemm2<Cfg, EpiTmaStore>(A, B, D, M, N, K, lda * 2, ldb * 2, ldd * 2, ep, 0, s);
}

// fp32 accumulate epilogue: D32 = beta * D32 + acc  (dH of the chunked linear-CE backward sums over vocab slices)
struct AccumF32Params {
  float* D;
  int64_t ldd;
  float beta;
};
struct EpiAccumF32 {
  using Params = AccumF32Params;
  __device__ __forceinline__ void begin(const Params&, int, int, const WorkInfo&) {}
  __device__ __forceinline__ void end(const Params&, int, int, const WorkInfo&) {}
  template <int BN>
  __device__ __forceinline__ void tile(const Params& p, const AccReader& acc, int row, int M, int n0, int N) {
#pragma unroll 1
    for (int c = 0; c < BN; c += 32) {
      float v[32];
      acc.load(c, v);
      const int col = n0 + c;
      if (row >= M || col >= N) continue;
      float* dst = p.D + static_cast<int64_t>(row) * p.ldd + col;
      const int valid = min(32, N - col);
      if (valid == 32) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 o = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
          if (p.beta != 0.f) {
            const float4 q = *reinterpret_cast<const float4*>(dst + 4 * i);
            o.x += p.beta * q.x; o.y += p.beta * q.y; o.z += p.beta * q.z; o.w += p.beta * q.w;
          }
          *reinterpret_cast<float4*>(dst + 4 * i) = o;
        }
      } else {
        for (int i = 0; i < valid; ++i) dst[i] = (p.beta != 0.f ? p.beta * dst[i] : 0.f) + v[i];
      }
    }
  }
};

template <bool A_MN, bool B_MN>
static int run2_accum(const void* A, const void* B, float* D, float beta, int M, int N, int K, int64_t lda, int64_t ldb,
                      int64_t ldd, cudaStream_t s) {
  using Cfg = Gemm2Cfg<256, A_MN, B_MN, false>;
  AccumF32Params ep{D, ldd, beta};
  return launch_gemm2<Cfg, EpiAccumF32>(A, B, nullptr, M, N, K, lda * 2, ldb * 2, 0, ep, 0, s);
}

// The 256x128 pair tile runs at about half the MMA rate of 256x256 (measured: profiles/r2_gemm2_microbench_v0.jsonl),
// so it is only used when the problem has no second 128-column half to fill.
static int pick_bn2(int /*M*/, int N) { return N <= 128 ? 128 : 256; }

}  // namespace prl

using namespace prl;

// D[M,N] = sum_k A(m,k) B(n,k) (+R).  Pitches in ELEMENTS.
//   a_mn = 0: A is [M,K] row-major;  a_mn = 1: A is [K,M] row-major (MN-major operand)
//   b_mn = 0: B is [N,K] row-major;  b_mn = 1: B is [K,N] row-major
extern "C" int prl_gemm2_bf16(const void* A, const void* B, void* D, const void* R, int M, int N, int K, int64_t lda,
                              int64_t ldb, int64_t ldd, int a_mn, int b_mn, int force_bn, cudaStream_t stream) {
  const int bn = force_bn > 0 ? force_bn : pick_bn2(M, N);
  if (bn != 128 && bn != 256) return -10;
#define PRL_G2(BN_, AM_, BM_) return run2<BN_, AM_, BM_>(A, B, D, R, M, N, K, lda, ldb, ldd, stream)
  if (bn == 256) {
    if (!a_mn && !b_mn) PRL_G2(256, false, false);
    if (!a_mn && b_mn) PRL_G2(256, false, true);
    if (a_mn && b_mn) PRL_G2(256, true, true);
    PRL_G2(256, true, false);
  } else {
    if (!a_mn && !b_mn) PRL_G2(128, false, false);
    if (!a_mn && b_mn) PRL_G2(128, false, true);
    if (a_mn && b_mn) PRL_G2(128, true, true);
    PRL_G2(128, true, false);
  }
#undef PRL_G2
}

// D32[M,N] = beta * D32 + sum_k A(m,k) B(n,k); fp32 output (ldd in elements, multiple of 4), layouts as above.
extern "C" int prl_gemm2_bf16_accum_f32(const void* A, const void* B, float* D, float beta, int M, int N, int K,
                                        int64_t lda, int64_t ldb, int64_t ldd, int a_mn, int b_mn, cudaStream_t stream) {
  if (!a_mn && !b_mn) return run2_accum<false, false>(A, B, D, beta, M, N, K, lda, ldb, ldd, stream);
  if (!a_mn && b_mn) return run2_accum<false, true>(A, B, D, beta, M, N, K, lda, ldb, ldd, stream);
  if (a_mn && b_mn) return run2_accum<true, true>(A, B, D, beta, M, N, K, lda, ldb, ldd, stream);
  return run2_accum<true, false>(A, B, D, beta, M, N, K, lda, ldb, ldd, stream);
}
