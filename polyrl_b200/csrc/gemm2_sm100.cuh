// polyrl_b200 — CTA-pair (cta_group::2) persistent tcgen05 GEMM family for sm_100a.
//
//   D[M,N] = sum_k A(m,k) * B(n,k)
//
// One mainloop serves the three layouts a Linear layer needs (reference: the cuBLAS calls behind
// nn.Linear forward/backward in 3rdparty/verl/verl/workers/actor/dp_actor.py:86-273 and the CE
// backward GEMMs of verl/utils/kernel/kernels.py:1571-1581):
//   NT  forward   y  = x  · W^T     A = x  [M,K]  K-major      B = W  [N,K]  K-major
//   NN  dgrad     dx = dy · W       A = dy [M,Kr] K-major      B = W  [Kr,N] MN-major
//   TN  wgrad     dW = dy^T · x     A = dy [Kr,M] MN-major     B = x  [Kr,N] MN-major
// (Kr = reduction length).  MN-major operands are staged by {64 x 64} SWIZZLE_128B TMA boxes and
// consumed through MN-major UMMA descriptors, so no operand is ever transposed in memory.
//
// A cluster of two CTAs (one SM pair) owns a 256 x BN output tile: each CTA stages its own 128
// rows of A and HALF of the B tile (BN/2 rows), the leader CTA's single elected thread issues
// `tcgen05.mma.cta_group::2` (M = 256) which reads both CTAs' shared memory, and each CTA's TMEM
// receives its 128 accumulator rows.  Halving the per-SM B traffic is what lifts the kernel off
// the shared-memory bandwidth limit of the 1-CTA 128x256 tile (gemm_sm100.cuh).
//
// Warp roles per CTA (192 threads): warp 0 TMA producer (both CTAs, completion bytes counted on the
// LEADER's full barrier), warp 1 TMEM allocator + (leader only) MMA issuer, warps 2..5 epilogue.
// Barriers: full[s] leader-only (count 1, expect 2 x stage bytes); empty[s] and tmem_full[a] per
// CTA, signalled by multicast `tcgen05.commit`; tmem_empty[a] on the leader, 8 remote arrivals
// (one elected lane per epilogue warp of both CTAs).
//
// Epilogues: (a) functor mode — same `Epi` interface as gemm_sm100.cuh (SwiGLU, linear-CE, ...);
// (b) TMA-store mode — accumulators -> bf16 -> 128B-swizzled shared staging -> `cp.async.bulk.tensor`
// store (UTMASTG), optional fused residual add; each epilogue warp double-buffers its own 32 x 64
// box so no cross-warp barrier is needed.
#pragma once
#include "gemm_sm100.cuh"

namespace prl {

constexpr int kSmemMax = 232448;   // 227 KB opt-in limit per CTA

template <int BN, bool A_MN_, bool B_MN_, bool TMA_STORE_>
struct Gemm2Cfg {
  static constexpr bool A_MN = A_MN_, B_MN = B_MN_, TMA_STORE = TMA_STORE_;
  static constexpr int BLOCK_M = 256;                 // per CTA pair
  static constexpr int CTA_M = 128;
  static constexpr int BLOCK_N = BN;
  static constexpr int CTA_N = BN / 2;                // rows of B staged by each CTA
  static constexpr int BLOCK_K = 64;
  static constexpr int A_BYTES = CTA_M * 128;
  static constexpr int B_BYTES = CTA_N * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int EPI_BYTES = TMA_STORE ? 4 * 2 * 4096 : 0;
  static constexpr int STAGES_RAW = (kSmemMax - 1024 - 512 - EPI_BYTES) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BN;            // 256 or 512
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 + 512;
  static_assert(BN == 128 || BN == 256, "CTA-pair tile is 256 x {128,256}");
  static_assert(STAGES >= 3, "pipeline too shallow");
};

// TMA-store epilogue tag (handled inside the kernel, not a per-row functor).
struct TmaStoreParams {
  const __nv_bfloat16* R;   // optional residual [M, ldr]
  int64_t ldr;
};
struct EpiTmaStore {
  using Params = TmaStoreParams;
};

template <class Cfg, class Epi>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kNumThreads, 1)
gemm2_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                   const __grid_constant__ CUtensorMap tmD, typename Epi::Params ep, int M, int N, int K,
                   int num_m_blocks, int num_n_blocks, int n_splits, int nb_per_split, Sched sc) {
  constexpr int BN = Cfg::BLOCK_N;
  constexpr int kStages = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::A_BYTES;
  uint8_t* smem_epi = smem + kStages * Cfg::STAGE_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_epi + Cfg::EPI_BYTES);
  uint64_t* full_bar = bars;                       // [kStages]  TMA (both CTAs) -> leader MMA
  uint64_t* empty_bar = bars + kStages;            // [kStages]  MMA -> TMA, one copy per CTA
  uint64_t* tmem_full = bars + 2 * kStages;        // [2]        MMA -> epilogue, one copy per CTA
  uint64_t* tmem_empty = bars + 2 * kStages + 2;   // [2]        epilogue (both CTAs) -> leader MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1;
  const int num_clusters = gridDim.x >> 1;
  const int total_work = num_m_blocks * n_splits;
  const int num_kb = (K + Cfg::BLOCK_K - 1) / Cfg::BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if constexpr (Cfg::TMA_STORE) tma_prefetch_desc(&tmD);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc_2sm(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish_2sm();
  }
  tc_fence_before();
  cluster_sync_all();            // barrier inits + TMEM allocation visible across the pair
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const WorkInfo wi = decode_work(w, num_m_blocks, num_n_blocks, n_splits, nb_per_split, sc);
        const int m0 = wi.m_blk * Cfg::BLOCK_M + static_cast<int>(rank) * Cfg::CTA_M;
        for (int nbs = wi.n_begin; nbs < wi.n_end; ++nbs) {
          const int n0 = rot_n(nbs, num_n_blocks, sc) * BN + static_cast<int>(rank) * Cfg::CTA_N;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            if (rank == 0) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
            uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
            uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
            const int k0 = kb * Cfg::BLOCK_K;
            if constexpr (!Cfg::A_MN) {
              tma_load_2d_2sm(&tmA, &full_bar[stage], sa, k0, m0, kEvictNormal);
            } else {
#pragma unroll
              for (int c = 0; c < Cfg::CTA_M / 64; ++c)
                tma_load_2d_2sm(&tmA, &full_bar[stage], sa + c * 8192, m0 + c * 64, k0, kEvictNormal);
            }
            if constexpr (!Cfg::B_MN) {
              tma_load_2d_2sm(&tmB, &full_bar[stage], sb, k0, n0, kEvictNormal);
            } else {
#pragma unroll
              for (int c = 0; c < Cfg::CTA_N / 64; ++c)
                tma_load_2d_2sm(&tmB, &full_bar[stage], sb + c * 8192, n0 + c * 64, k0, kEvictNormal);
            }
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ============================
    if (rank == 0 && lane == 0) {
      constexpr uint32_t idesc = umma_idesc_major(Cfg::BLOCK_M, BN, 1, 1, Cfg::A_MN ? 1 : 0, Cfg::B_MN ? 1 : 0);
      constexpr uint32_t a_step = Cfg::A_MN ? (16 * 128) >> 4 : 32 >> 4;   // descriptor units (16 B) per UMMA_K
      constexpr uint32_t b_step = Cfg::B_MN ? (16 * 128) >> 4 : 32 >> 4;
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_iter = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const WorkInfo wi = decode_work(w, num_m_blocks, num_n_blocks, n_splits, nb_per_split, sc);
        for (int nb = wi.n_begin; nb < wi.n_end; ++nb, ++tile_iter) {
          const uint32_t as = tile_iter & 1;
          const uint32_t aphase = (tile_iter >> 1) & 1;
          mbar_wait(&tmem_empty[as], aphase ^ 1);
          tc_fence_after();
          const uint32_t tmem_d = tmem_base + as * BN;
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&full_bar[stage], phase);
            tc_fence_after();
            const uint8_t* sa = smem_a + stage * Cfg::A_BYTES;
            const uint8_t* sb = smem_b + stage * Cfg::B_BYTES;
            const uint64_t da = Cfg::A_MN ? umma_desc_sw128_mnmajor(sa, 8192) : umma_desc_sw128_kmajor(sa);
            const uint64_t db = Cfg::B_MN ? umma_desc_sw128_mnmajor(sb, 8192) : umma_desc_sw128_kmajor(sb);
#pragma unroll
            for (int k = 0; k < Cfg::BLOCK_K / 16; ++k) {
              const uint32_t acc = (kb | k) != 0;
              umma_bf16_ss_2sm(tmem_d, da + k * a_step, db + k * b_step, idesc, acc);
            }
            umma_commit_2sm(&empty_bar[stage], 3);      // both CTAs' producers may refill this slot
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit_2sm(&tmem_full[as], 3);           // both CTAs' epilogues may drain their half
        }
      }
    }
  } else {
    // ================================ epilogue (both CTAs) ====================================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
    const int row_in_tile = static_cast<int>(rank) * Cfg::CTA_M + quad * 32 + lane;
    uint32_t tile_iter = 0;
    if constexpr (Cfg::TMA_STORE) {
      uint8_t* ebuf = smem_epi + (warp - 2) * 8192;
      uint32_t chunk_ctr = 0;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const WorkInfo wi = decode_work(w, num_m_blocks, num_n_blocks, n_splits, nb_per_split, sc);
        const int row = wi.m_blk * Cfg::BLOCK_M + row_in_tile;
        const int row0 = row - lane;
        for (int nb = wi.n_begin; nb < wi.n_end; ++nb, ++tile_iter) {
          const uint32_t as = tile_iter & 1;
          const uint32_t aphase = (tile_iter >> 1) & 1;
          const int n0 = rot_n(nb, num_n_blocks, sc) * BN;
          mbar_wait(&tmem_full[as], aphase);
          tc_fence_after();
          const uint32_t taddr = tmem_base + as * BN + (static_cast<uint32_t>(quad * 32) << 16);
#pragma unroll 1
          for (int c = 0; c < BN / 64; ++c, ++chunk_ctr) {
            uint8_t* buf = ebuf + (chunk_ctr & 1) * 4096;
            if (lane == 0) tma_store_wait_read<1>();   // the store that last read `buf` has drained it
            __syncwarp();
            uint32_t r[64];
            tmem_ld_32x32(taddr + c * 64, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
            tmem_ld_32x32(taddr + c * 64 + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
            tmem_ld_wait();
            if (c == BN / 64 - 1) {                     // accumulator fully read: hand TMEM back early
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_cluster(&tmem_empty[as], 0);
            }
            const int col0 = n0 + c * 64;
            if (ep.R != nullptr && row < M) {
              const __nv_bfloat16* rp = ep.R + static_cast<int64_t>(row) * ep.ldr + col0;
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                if (col0 + 8 * j < N) {
                  const int4 q = ld_nc_v4(rp + 8 * j);
                  const float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), cc = unpack_bf16(q.z), d = unpack_bf16(q.w);
                  r[8 * j + 0] = __float_as_uint(__uint_as_float(r[8 * j + 0]) + a.x);
                  r[8 * j + 1] = __float_as_uint(__uint_as_float(r[8 * j + 1]) + a.y);
                  r[8 * j + 2] = __float_as_uint(__uint_as_float(r[8 * j + 2]) + b.x);
                  r[8 * j + 3] = __float_as_uint(__uint_as_float(r[8 * j + 3]) + b.y);
                  r[8 * j + 4] = __float_as_uint(__uint_as_float(r[8 * j + 4]) + cc.x);
                  r[8 * j + 5] = __float_as_uint(__uint_as_float(r[8 * j + 5]) + cc.y);
                  r[8 * j + 6] = __float_as_uint(__uint_as_float(r[8 * j + 6]) + d.x);
                  r[8 * j + 7] = __float_as_uint(__uint_as_float(r[8 * j + 7]) + d.y);
                }
              }
            }
            // row `lane` of the 32 x 64 box: 8 x 16 B chunks, chunk j lands at (j ^ (lane & 7)) (128B swizzle)
            const uint32_t row_base = smem_u32(buf) + lane * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const uint32_t p0 = pack_bf16(__uint_as_float(r[8 * j + 0]), __uint_as_float(r[8 * j + 1]));
              const uint32_t p1 = pack_bf16(__uint_as_float(r[8 * j + 2]), __uint_as_float(r[8 * j + 3]));
              const uint32_t p2 = pack_bf16(__uint_as_float(r[8 * j + 4]), __uint_as_float(r[8 * j + 5]));
              const uint32_t p3 = pack_bf16(__uint_as_float(r[8 * j + 6]), __uint_as_float(r[8 * j + 7]));
              asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(row_base + ((j ^ (lane & 7)) << 4)),
                           "r"(p0), "r"(p1), "r"(p2), "r"(p3)
                           : "memory");
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) {
              if (row0 < M && col0 < N) tma_store_2d(&tmD, buf, col0, row0);   // TMA clips partial boxes
              tma_store_commit();
            }
          }
        }
      }
      if (lane == 0) tma_store_wait<0>();
    } else {
      Epi epi;
      for (int w = cluster_id; w < total_work; w += num_clusters) {
        const WorkInfo wi = decode_work(w, num_m_blocks, num_n_blocks, n_splits, nb_per_split, sc);
        const int row = wi.m_blk * Cfg::BLOCK_M + row_in_tile;
        epi.begin(ep, row, M, wi);
        for (int nb = wi.n_begin; nb < wi.n_end; ++nb, ++tile_iter) {
          const uint32_t as = tile_iter & 1;
          const uint32_t aphase = (tile_iter >> 1) & 1;
          mbar_wait(&tmem_full[as], aphase);
          tc_fence_after();
          AccReader acc{tmem_base + as * BN + (static_cast<uint32_t>(quad * 32) << 16)};
          epi.template tile<BN>(ep, acc, row, M, rot_n(nb, num_n_blocks, sc) * BN, N);
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(&tmem_empty[as], 0);
        }
        epi.end(ep, row, M, wi);
      }
    }
  }

  tc_fence_before();
  cluster_sync_all();            // nobody leaves while the peer may still read its smem / signal its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host launcher.  Pitches in BYTES.  For an MN-major operand the pointer addresses the
// [reduction, M-or-N] row-major matrix (i.e. the tensor as it lies in memory).
// ------------------------------------------------------------------------------------------
template <class Cfg, class Epi>
int launch_gemm2(const void* A, const void* B, void* D, int M, int N, int K, int64_t lda_bytes, int64_t ldb_bytes,
                 int64_t ldd_bytes, const typename Epi::Params& ep, int n_splits, cudaStream_t stream,
                 Sched sc = Sched{0, 1, 0}) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  CUtensorMap tmA, tmB, tmD;
  if (Cfg::A_MN ? make_tmap_kmajor(&tmA, A, K, M, lda_bytes, 64, 2) : make_tmap_kmajor(&tmA, A, M, K, lda_bytes, Cfg::CTA_M, 2))
    return -2;
  if (Cfg::B_MN ? make_tmap_kmajor(&tmB, B, K, N, ldb_bytes, 64, 2) : make_tmap_kmajor(&tmB, B, N, K, ldb_bytes, Cfg::CTA_N, 2))
    return -3;
  if (Cfg::TMA_STORE) {
    if (make_tmap_kmajor(&tmD, D, M, N, ldd_bytes, 32, 2)) return -4;
  } else {
    tmD = tmA;
  }
  const int num_m_blocks = (M + Cfg::BLOCK_M - 1) / Cfg::BLOCK_M;
  const int num_n_blocks = (N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N;
  if (n_splits <= 0 || n_splits > num_n_blocks) n_splits = num_n_blocks;
  const int nb_per_split = (num_n_blocks + n_splits - 1) / n_splits;
  n_splits = (num_n_blocks + nb_per_split - 1) / nb_per_split;
  auto kern = gemm2_sm100_kernel<Cfg, Epi>;
  static bool attr_set = false;
  if (!attr_set) {
    PRL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  int clusters = sm_count() / 2;
  const int total = num_m_blocks * n_splits;
  if (total < clusters) clusters = total;
  kern<<<2 * clusters, kNumThreads, Cfg::SMEM_BYTES, stream>>>(tmA, tmB, tmD, ep, M, N, K, num_m_blocks, num_n_blocks,
                                                               n_splits, nb_per_split, sc);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace prl
