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
