// polyrl_b200 — decode-shape tcgen05 GEMM v2 ("stream GEMM"):  out[B, N] = X[B, K] · W[N, K]^T,  B <= 256.
//
// What bounds a decode GEMM on B200 (measured, profiles/r2_decode_gemm_experiments.md): at B = 128 the
// tensor work is 37 % of peak at HBM speed and a 128x128x16 SS-mode `tcgen05.mma` costs ~140 clk whatever
// the batch (the A-operand read from shared memory is exposed), so one SM can turn at most ~29 B/clk of
// weights into MMAs while HBM delivers ~23 B/clk per SM.  A kernel that leaves SMs idle is therefore MMA
// bound, not memory bound (round 1's 112-CTA gate/up kernel: 54 us against a 36 us HBM floor), and the
// activation multicast tried first changed nothing.  The design follows from that:
//
//   * swap-AB: UMMA M = 128 weight rows, UMMA N = padded batch; every CTA streams its own weight panel
//     exactly once through a 6..8 stage TMA ring (evict-first), activations come from L2 (evict-last).
//   * split-K over a thread-block CLUSTER: the S CTAs of a cluster own the same 128 weight rows and 1/S of
//     K each, so tiles x S CTAs cover (almost) every SM.  The partial accumulators never touch global
//     memory: after the mainloop each CTA scatters its TMEM tile in 8-column chunks into the shared memory
//     of the chunk's owner CTA (st.shared::cluster over DSMEM, the idle TMA ring is the inbox), one
//     cluster barrier later every CTA reduces the chunks it owns and runs the fused epilogue for them:
//     a reduce-scatter with no workspace, no counters and no fences.
//   * SwiGLU tiles are 64 gate rows + 64 up rows of the block-interleaved weight (two 64-row TMA boxes), so
//     the gate/up GEMM has 224 tiles instead of 112 and the activation is applied when the owner reads its
//     chunk back from shared memory.
//   * epilogue: + bias | + residual | fp8 row/col scales | SwiGLU, 16-byte coalesced bf16 stores.
//
// Replaces the cuBLAS / sgl-kernel decode GEMMs of the reference stack (SURVEY §2.5b "Decode GEMMs";
// 3rdparty/verl/verl/workers/rollout/sglang_rollout/sglang_rollout.py drives SGLang's library kernels).
#include <cuda_runtime.h>
#include <stdlib.h>

#include "common.cuh"

namespace prl {

constexpr int kSgThreads = 192;          // warp0 TMA, warp1 MMA, warps 2-5 epilogue
constexpr int kSgSmemBudget = 196 * 1024;

struct StreamParams {
  __nv_bfloat16* out;          // [B, ldo]
  const __nv_bfloat16* res;    // optional residual [B, ldo]
  const __nv_bfloat16* bias;   // optional [N]
  const float* x_scale;        // optional fp8: per batch row
  const float* w_scale;        // optional fp8: per weight row
  int64_t ldo;
  int B, N, K, BP, S, kb_per_split, swiglu, tiles, dbg, nacc, pf;
};

__device__ __forceinline__ void st_shared_cluster_f32(uint32_t addr, float v) {
  asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}

// PAIR = true: the CTAs of a cluster are organised as S CTA PAIRS (cta_group::2).  A pair owns 256 weight rows (128 per CTA)
// and one k-range; each CTA stages half of the activation k-block and the leader issues ONE tcgen05.mma (M = 256) per K=16
// step for both — 256 weight rows per ~128-clk instruction instead of 128 rows per ~190 clk (profiles/r2_decode_gemm_
// experiments.md), which is what lifts the per-SM weight rate above the HBM share.
template <bool FP8, bool PAIR>
__global__ void __launch_bounds__(kSgThreads, 1)
stream_gemm_kernel(const __grid_constant__ CUtensorMap tmW, const __grid_constant__ CUtensorMap tmX, StreamParams p) {
  constexpr int BLOCK_K = FP8 ? 128 : 64;
  constexpr int A_BYTES = 128 * 128;
  constexpr int NC = PAIR ? 2 : 1;                     // CTAs that share one accumulator tile / MMA
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int b_rows = p.BP / NC;                        // activation rows staged by THIS CTA
  const int b_bytes = b_rows * 128;
  const int stage_bytes = A_BYTES + b_bytes;
  int stages = kSgSmemBudget / stage_bytes;
  if (stages > 8) stages = 8;
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + stages * A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kSgSmemBudget);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + 8;
  uint64_t* acc_bar = bars + 16;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 17);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int S = p.S;
  const int crank = static_cast<int>(blockIdx.x) % (S * NC);         // rank inside the cluster
  const int split = crank / NC;                                       // k-range owned by this CTA (pair)
  const int r2 = crank % NC;                                          // CTA inside the pair (0 = MMA leader)
  const int tile = static_cast<int>(blockIdx.x) / (S * NC);           // 128 (single) or 256 (pair) weight rows
  const int num_kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
  const int kb0 = split * p.kb_per_split;
  const int kb1 = min(num_kb_total, kb0 + p.kb_per_split);
  const int nkb = max(0, kb1 - kb0);
  // weight rows of this CTA's 128 TMEM lanes: lanes [0,64) <- row_g.., lanes [64,128) <- row_u..
  //   plain:  128 consecutive rows.   SwiGLU: 64 gate rows + the 64 matching up rows of the [128 gate | 128 up] block layout
  int row_g, row_u, n_tile0;
  if (p.swiglu) {
    const int t64 = tile * NC + r2;                                   // 64-feature tile index
    row_g = (t64 >> 1) * 256 + (t64 & 1) * 64;
    row_u = row_g + 128;
    n_tile0 = t64 * 64;
  } else {
    row_g = (tile * NC + r2) * 128;
    row_u = row_g + 64;
    n_tile0 = row_g;
  }
  const int NACC = PAIR ? 1 : p.nacc;
  uint32_t tmem_cols = 32;
  while (tmem_cols < static_cast<uint32_t>(p.BP * NACC)) tmem_cols <<= 1;
  const uint16_t pair_mask = static_cast<uint16_t>(3u << (2 * split));

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmW);
    tma_prefetch_desc(&tmX);
    for (int i = 0; i < stages; ++i) { mbar_init(&full_bar[i], 1); mbar_init(&empty_bar[i], 1); }
    mbar_init(acc_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) {
    if constexpr (PAIR) { tmem_alloc_2sm(tmem_slot, tmem_cols); tmem_relinquish_2sm(); }
    else { tmem_alloc(tmem_slot, tmem_cols); tmem_relinquish(); }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      // L2 prefetch `pf` k-blocks ahead of the ring: HBM latency under load (~2-3 us) x 23 B/clk per SM is more than the
      // ~128 KB of weights the ring can hold, so the ring alone cannot keep the HBM share of one SM in flight
      const int pf = p.pf;
      for (int kb = kb0; kb < min(kb1, kb0 + pf); ++kb) {
        tma_prefetch_l2_2d(&tmW, kb * BLOCK_K, row_g);
        tma_prefetch_l2_2d(&tmW, kb * BLOCK_K, row_u);
      }
      for (int kb = kb0; kb < kb1; ++kb) {
        if ((p.dbg & 2) && kb - kb0 >= stages) break;              // experiment: MMA-only (ring loaded once)
        if (kb + pf < kb1) {
          tma_prefetch_l2_2d(&tmW, (kb + pf) * BLOCK_K, row_g);
          tma_prefetch_l2_2d(&tmW, (kb + pf) * BLOCK_K, row_u);
        }
        mbar_wait(&empty_bar[stage], phase ^ 1);
        uint8_t* sa = smem_a + stage * A_BYTES;
        uint8_t* sb = smem_b + stage * b_bytes;
        if constexpr (PAIR) {
          if (r2 == 0) mbar_expect_tx(&full_bar[stage], 2 * stage_bytes);      // both CTAs' bytes are counted on the leader
          tma_load_2d_2sm(&tmW, &full_bar[stage], sa, kb * BLOCK_K, row_g, kEvictFirst);
          tma_load_2d_2sm(&tmW, &full_bar[stage], sa + 64 * 128, kb * BLOCK_K, row_u, kEvictFirst);
          tma_load_2d_2sm(&tmX, &full_bar[stage], sb, kb * BLOCK_K, r2 * b_rows, kEvictLast);
        } else {
          mbar_expect_tx(&full_bar[stage], stage_bytes);
          // weights are read exactly once: do not let them displace the activations in L2
          tma_load_2d_hint(&tmW, &full_bar[stage], sa, kb * BLOCK_K, row_g, kEvictFirst);
          tma_load_2d_hint(&tmW, &full_bar[stage], sa + 64 * 128, kb * BLOCK_K, row_u, kEvictFirst);
          tma_load_2d_hint(&tmX, &full_bar[stage], sb, kb * BLOCK_K, 0, kEvictLast);
        }
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && r2 == 0) {
      const uint32_t idesc = FP8 ? umma_idesc(128 * NC, p.BP, 0, 0) : umma_idesc(128 * NC, p.BP, 1, 1);
      int stage = 0; uint32_t phase = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        if (!((p.dbg & 2) && kb >= stages)) mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (!PAIR && (p.dbg & 1)) {                                // experiment: stream-only (no MMA)
          mbar_arrive(&empty_bar[stage]);
          if (++stage == stages) { stage = 0; phase ^= 1; }
          continue;
        }
        const uint64_t da = umma_desc_sw128_kmajor(smem_a + stage * A_BYTES);
        const uint64_t db = umma_desc_sw128_kmajor(smem_b + stage * b_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const int a = k & (NACC - 1);
          const uint32_t acc = (kb != 0) || (k >= NACC);
          if constexpr (PAIR) {
            umma_bf16_ss_2sm(tmem_base, da + 2 * k, db + 2 * k, idesc, acc);
          } else {
            if constexpr (FP8) umma_f8_ss(tmem_base + a * p.BP, da + 2 * k, db + 2 * k, idesc, acc);
            else umma_bf16_ss(tmem_base + a * p.BP, da + 2 * k, db + 2 * k, idesc, acc);
          }
        }
        if constexpr (PAIR) umma_commit_2sm(&empty_bar[stage], pair_mask);
        else if (!(p.dbg & 2)) umma_commit(&empty_bar[stage]);
        if (++stage == stages) { stage = 0; phase ^= 1; }
      }
      if constexpr (PAIR) umma_commit_2sm(acc_bar, pair_mask);
      else if (p.dbg & 1) mbar_arrive(acc_bar);
      else umma_commit(acc_bar);
    }
  }

  // ---------------------------------------------------------------- reduce-scatter over the cluster + epilogue
  // Batch columns are dealt to the S CTAs in chunks of 8 (chunk q -> CTA q % S, local slot q / S).  Inbox of a
  // CTA: [src S][local chunk LC][8 columns][128 rows] fp32, living in the TMA ring once every CTA of the
  // cluster is past its mainloop.
  const int num_chunks = p.BP >> 3;
  const int LC = (num_chunks + S - 1) / S;
  float* inbox = reinterpret_cast<float*>(smem);
  const bool epi_warp = warp >= 2;
  const int quad = warp & 3;
  const int r = quad * 32 + lane;                                // accumulator row (TMEM lane) of this thread
  if (epi_warp) {
    if (nkb > 0) { mbar_wait(acc_bar, 0); tc_fence_after(); }
  }
  if (S * NC > 1) cluster_sync_all();                            // every ring of the cluster is idle from here on
  else __syncthreads();
  if (epi_warp) {
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t inbox_u32 = smem_u32(inbox);
    const float ws_row = p.w_scale ? p.w_scale[min(p.N - 1, (r < 64 ? row_g : row_u - 64) + r)] : 1.f;
    for (int c = 0; c < p.BP; c += 32) {
      uint32_t v[32];
      __syncwarp();
      if (nkb > 0) {
        tmem_ld_32x32(tmem_base + lane_base + c, v);
        tmem_ld_wait();
        for (int a = 1; a < NACC; ++a) {
          uint32_t v2[32];
          tmem_ld_32x32(tmem_base + lane_base + a * p.BP + c, v2);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) + __uint_as_float(v2[i]));
        }
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = 0;
      }
#pragma unroll
      for (int qq = 0; qq < 4; ++qq) {
        const int q = (c >> 3) + qq;
        if (q >= num_chunks) break;
        const int owner = q % S, slot = q / S;
        const uint32_t local = inbox_u32 + static_cast<uint32_t>((((split * LC + slot) * 8) * 128 + r) * 4);
        const uint32_t dst = S > 1 ? mapa_u32(local, owner * NC + r2) : local;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          st_shared_cluster_f32(dst + i * 512, __uint_as_float(v[qq * 8 + i]) * ws_row);   // 32 lanes -> 128 contiguous bytes
      }
    }
  }
  tc_fence_before();
  if (S * NC > 1) cluster_sync_all();                            // all partial chunks have landed in their owners
  else __syncthreads();
  if (epi_warp) {
    const int et = (warp - 2) * 32 + lane;                       // 0..127
    const int feats = p.swiglu ? 64 : 128;                       // output features of this CTA
    const int n_out = p.swiglu ? p.N / 2 : p.N;
    const int segs = feats >> 3;
    const int seg = et % segs, bsub = et / segs;
    const int bstep = 128 / segs;                                // batch-column slots covered per pass
    const int n0 = n_tile0 + seg * 8;
    if (n0 < n_out) {
      float bv[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) bv[q] = (p.bias && !p.swiglu) ? __bfloat162float(p.bias[n0 + q]) : 0.f;
      const int my_chunks = (num_chunks - split + S - 1) / S;    // chunks q = split, split + S, ...
      for (int lc = bsub; lc < my_chunks * 8; lc += bstep) {
        const int slot = lc >> 3, i = lc & 7;
        const int b = (slot * S + split) * 8 + i;                // batch row
        if (b >= p.B) continue;
        float y[8], u[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) { y[q] = 0.f; u[q] = 0.f; }
        for (int src = 0; src < S; ++src) {
          const float* row = inbox + (((src * LC + slot) * 8 + i) * 128);
          const float4 lo = *reinterpret_cast<const float4*>(row + seg * 8);
          const float4 hi = *reinterpret_cast<const float4*>(row + seg * 8 + 4);
          y[0] += lo.x; y[1] += lo.y; y[2] += lo.z; y[3] += lo.w; y[4] += hi.x; y[5] += hi.y; y[6] += hi.z; y[7] += hi.w;
          if (p.swiglu) {
            const float4 ulo = *reinterpret_cast<const float4*>(row + 64 + seg * 8);
            const float4 uhi = *reinterpret_cast<const float4*>(row + 64 + seg * 8 + 4);
            u[0] += ulo.x; u[1] += ulo.y; u[2] += ulo.z; u[3] += ulo.w; u[4] += uhi.x; u[5] += uhi.y; u[6] += uhi.z; u[7] += uhi.w;
          }
        }
        const float xs = p.x_scale ? p.x_scale[b] : 1.f;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          if (p.swiglu) {
            const float g = y[q] * xs, uu = u[q] * xs;
            y[q] = g / (1.f + __expf(-g)) * uu;
          } else {
            y[q] = y[q] * xs + bv[q];
          }
        }
        const int64_t o = static_cast<int64_t>(b) * p.ldo + n0;
        if (p.res) {
          const int4 q4 = ld_nc_v4(p.res + o);
          const float2 r0 = unpack_bf16(q4.x), r1 = unpack_bf16(q4.y), r2 = unpack_bf16(q4.z), r3 = unpack_bf16(q4.w);
          y[0] += r0.x; y[1] += r0.y; y[2] += r1.x; y[3] += r1.y; y[4] += r2.x; y[5] += r2.y; y[6] += r3.x; y[7] += r3.y;
        }
        int4 ov;
        ov.x = pack_bf16(y[0], y[1]); ov.y = pack_bf16(y[2], y[3]); ov.z = pack_bf16(y[4], y[5]); ov.w = pack_bf16(y[6], y[7]);
        st_v4(p.out + o, ov);
      }
    }
  }
  tc_fence_before();
  if constexpr (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if constexpr (PAIR) tmem_dealloc_2sm(tmem_base, tmem_cols); else tmem_dealloc(tmem_base, tmem_cols);
  }
}

using EncodeTiledFn3 = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn3 sg_encode_fn() {
  static EncodeTiledFn3 fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) != cudaSuccess || !q) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn3>(q);
  }
  return fn;
}
static int sg_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_bytes, uint32_t box_rows, int eb) {
  ensure_thread_ctx();
  auto fn = sg_encode_fn();
  if (!fn) return -1;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / eb), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, eb == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), gdim,
                  gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r);
}

template <bool FP8, bool PAIR>
static int launch_stream(const void* X, const void* W, StreamParams p, int64_t ldx_bytes, int64_t ldw_bytes, cudaStream_t s) {
  constexpr int EB = FP8 ? 1 : 2;
  constexpr int NC = PAIR ? 2 : 1;
  CUtensorMap tmW, tmX;
  if (sg_tmap(&tmW, W, p.N, p.K, ldw_bytes, 64, EB)) return -61;       // two 64-row boxes per stage (gate | up halves)
  if (sg_tmap(&tmX, X, p.B, p.K, ldx_bytes, p.BP / NC, EB)) return -62;
  auto kern = stream_gemm_kernel<FP8, PAIR>;
  static bool attr = false;
  const int smem = kSgSmemBudget + 1024 + 256;
  if (!attr) {
    PRL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    attr = true;
  }
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3(p.tiles * p.S * NC);
  cfg.blockDim = dim3(kSgThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = p.S * NC;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  PRL_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmW, tmX, p));
  return 0;
}

}  // namespace prl

using namespace prl;

// mode: 0 = store (+bias, +residual), 1 = SwiGLU (W block-interleaved [128 gate | 128 up] per 256 rows, out is [B, N/2]).
// dtype: 0 = bf16, 1 = fp8 e4m3 (x_scale per batch row, w_scale per weight row required).
// splits: k-splits S per cluster (0 = auto).  pair: 1 = CTA pairs (cta_group::2, 256 weight rows per MMA, bf16 only),
// 0 = single CTAs, -1 = auto (pairs whenever the shape allows).
extern "C" int prl_stream_gemm(const void* X, const void* W, void* out, const void* res, const void* bias, const float* x_scale,
                               const float* w_scale, int B, int N, int K, int64_t ldx, int64_t ldw, int64_t ldo, int mode,
                               int dtype, int splits, int pair, cudaStream_t stream) {
  if (B <= 0 || B > 256) return -60;
  StreamParams p;
  {
    const char* e = getenv("PRL_STREAM_DBG");                       // kernel experiments only (scripts/dbg_stream_gemm.py)
    p.dbg = e ? atoi(e) : 0;
  }
  p.out = static_cast<__nv_bfloat16*>(out);
  p.res = static_cast<const __nv_bfloat16*>(res);
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.x_scale = x_scale; p.w_scale = w_scale;
  p.ldo = ldo; p.B = B; p.N = N; p.K = K;
  p.BP = ((B + 15) / 16) * 16;
  p.swiglu = mode == 1;
  if (mode == 1 && N % 256 != 0) return -63;
  const int n_out = mode == 1 ? N / 2 : N;
  if (n_out % 8 != 0 || ldo % 8 != 0) return -65;
  const int block_k = dtype == 1 ? 128 : 64;
  const int num_kb = (K + block_k - 1) / block_k;
  const int tiles128 = (N + 127) / 128;
  bool use_pair = pair == 1 || (pair < 0 && dtype == 0 && tiles128 % 2 == 0 && p.BP >= 32);
  if (use_pair && (dtype != 0 || tiles128 % 2 != 0)) return -68;
  const int NC = use_pair ? 2 : 1;
  p.tiles = tiles128 / NC;
  int S = splits;
  if (S <= 0) {
    // one wave of clusters (a cluster needs all its SMs inside one GPC: count ~132 usable SMs beyond 2-CTA clusters), every
    // split keeps >= 8 k-blocks
    S = 1;
    for (int s = 2; s * NC <= 8; ++s) {
      const int usable = s * NC <= 2 ? 148 : 132;
      if (p.tiles * s * NC <= usable && num_kb / s >= 8) S = s;
    }
  }
  if (S * NC > 8 || S > num_kb) return -66;
  p.kb_per_split = (num_kb + S - 1) / S;
  if ((S - 1) * p.kb_per_split >= num_kb) return -67;              // every rank of the cluster must own >= 1 k-block
  p.S = S;
  {
    const char* e = getenv("PRL_STREAM_NACC");
    int na = e ? atoi(e) : 1;
    while (na > 1 && na * p.BP > 512) na >>= 1;
    p.nacc = (na == 1 || na == 2 || na == 4) ? na : 1;
  }
  {
    const char* e = getenv("PRL_STREAM_PF");                        // L2 prefetch distance in k-blocks (0 = off)
    p.pf = e ? atoi(e) : 0;                        // measured: prefetching only adds traffic (profiles/r2_decode_gemm_experiments.md)
  }
  const int eb = dtype == 1 ? 1 : 2;
  if (use_pair) return launch_stream<false, true>(X, W, p, ldx * eb, ldw * eb, stream);
  return dtype == 1 ? launch_stream<true, false>(X, W, p, ldx * eb, ldw * eb, stream)
                    : launch_stream<false, false>(X, W, p, ldx * eb, ldw * eb, stream);
}
