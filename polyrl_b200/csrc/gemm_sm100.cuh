// polyrl_b200 — persistent warp-specialised tcgen05 GEMM mainloop for sm_100a.
//
//   D[M,N] = A[M,K] · B[N,K]^T          (both operands K-major, i.e. nn.Linear layout)
//
// One CTA per SM, 192 threads:
//   warp 0      TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B boxes, kStages-deep ring)
//   warp 1      MMA issuer     (one thread issues tcgen05.mma 128xBNx{16|32}; accumulators in TMEM)
//   warps 2..5  epilogue       (tcgen05.ld TMEM -> registers -> fused epilogue functor -> HBM)
// Two TMEM accumulator stages (2*BN columns) let the epilogue of tile i overlap the mainloop of
// tile i+1.  The epilogue is a functor (`Epi`) so the same mainloop serves: plain store,
// residual add, SwiGLU, fp8 row/col scaling, the fused linear-cross-entropy statistics pass and
// its d_logits pass (linear_ce.cu), and the all-gather-fused variant (ag_gemm.cu) which gates the
// producer on per-shard arrival flags written by peer-copy CTAs of the same kernel.
//
// Work decomposition: a work item is (m_block, n_split); a split owns a contiguous range of
// n-blocks that the CTA walks while the epilogue functor carries per-row running state
// (needed by the online-softmax epilogue).  Plain GEMMs use one n-block per split.
#pragma once
#include "common.cuh"

namespace prl {

constexpr int kBlockM = 128;
constexpr int kNumThreads = 192;
constexpr int kGroupM = 8;
constexpr int kSmemBudget = 200 * 1024;

template <int BN, int ELEM_BYTES>
struct GemmCfg {
  static constexpr int BLOCK_N = BN;
  static constexpr int BLOCK_K = 128 / ELEM_BYTES;        // one 128-byte swizzle row of K per stage
  static constexpr int UMMA_K = 32 / ELEM_BYTES;          // 16 for bf16, 32 for fp8
  static constexpr int A_BYTES = kBlockM * 128;
  static constexpr int B_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES_RAW = kSmemBudget / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int TMEM_COLS = 2 * BN < 32 ? 32 : 2 * BN;   // power of two for BN in {32,64,128,256}
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct WorkInfo {
  int m_blk, n_begin, n_end, split;
};

// Tile rasterisation.  raster 0: groups of kGroupM m-blocks x all n-splits (weights stay in L2 while
// activations stream).  raster 1: bands of `band_n` n-splits x all m-blocks, n advancing slowly —
// used by the all-gather-fused GEMM so tiles consume weight shards in their arrival order.
// n_rot rotates the n-block order so every rank starts with the shard it already owns.
struct Sched {
  int raster;
  int band_n;
  int n_rot;
};

__device__ __forceinline__ WorkInfo decode_work(int w, int num_m_blocks, int num_n_blocks, int n_splits,
                                                int nb_per_split, const Sched& sc) {
  WorkInfo wi;
  if (sc.raster == 0) {
    const int group_size = kGroupM * n_splits;
    const int g = w / group_size;
    const int first_m = g * kGroupM;
    const int gm = min(kGroupM, num_m_blocks - first_m);
    const int r = w - g * group_size;
    wi.m_blk = first_m + r % gm;
    wi.split = r / gm;
  } else {
    const int band_size = sc.band_n * num_m_blocks;
    const int b = w / band_size;
    const int first_n = b * sc.band_n;
    const int bn = min(sc.band_n, n_splits - first_n);
    const int r = w - b * band_size;
    wi.split = first_n + r % bn;
    wi.m_blk = r / bn;
  }
  wi.n_begin = wi.split * nb_per_split;
  wi.n_end = min(wi.n_begin + nb_per_split, num_n_blocks);
  return wi;
}
__device__ __forceinline__ int rot_n(int nb, int num_n_blocks, const Sched& sc) {
  const int r = nb + sc.n_rot;
  return r >= num_n_blocks ? r - num_n_blocks : r;
}

// Accessor handed to epilogue functors: loads 32 consecutive fp32 accumulator columns of this
// thread's row.  Must be called warp-uniformly (tcgen05.ld is .sync.aligned).
struct AccReader {
  uint32_t taddr;   // TMEM address of (lane quadrant base, accumulator stage column 0)
  __device__ __forceinline__ void load(int col, float (&v)[32]) const {
    uint32_t r[32];
    __syncwarp();
    tmem_ld_32x32(taddr + static_cast<uint32_t>(col), r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
  }
};

// Communication policy.  kCommWarps extra warps per CTA run `Comm::run` (peer-memory copies) while
// the TMA producer consults `Comm::wait_n_block` before touching the rows of B of an n-block.
struct NoComm {
  struct Params {};
  static constexpr int kCommWarps = 0;
  __device__ __forceinline__ static void wait_n_block(const Params&, int /*n_blk*/) {}
  __device__ __forceinline__ static void run(const Params&, int /*tid*/) {}
};
using NoGate = NoComm;

template <class Cfg, class Epi, class Gate, bool IS_FP8>
__global__ void __launch_bounds__(kNumThreads + 32 * Gate::kCommWarps, 1)
gemm_sm100_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                  typename Epi::Params ep, typename Gate::Params gp, int M, int N, int K, int num_m_blocks,
                  int num_n_blocks, int n_splits, int nb_per_split, Sched sc) {
  constexpr int BN = Cfg::BLOCK_N;
  constexpr int kStages = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + kStages * Cfg::A_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kStages * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;                       // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;            // [kStages]  MMA -> TMA
  uint64_t* tmem_full = bars + 2 * kStages;        // [2]        MMA -> epilogue
  uint64_t* tmem_empty = bars + 2 * kStages + 2;   // [2]        epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_work = num_m_blocks * n_splits;
  const int num_kb = (K + Cfg::BLOCK_K - 1) / Cfg::BLOCK_K;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        const WorkInfo wi = decode_work(w, num_m_blocks, num_n_blocks, n_splits, nb_per_split, sc);
        for (int nbs = wi.n_begin; nbs < wi.n_end; ++nbs) {
          const int nb = rot_n(nbs, num_n_blocks, sc);
          Gate::wait_n_block(gp, nb);
          for (int kb = 0; kb < num_kb; ++kb) {
            mbar_wait(&empty_bar[stage], phase ^ 1);
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(&tmA, &full_bar[stage], smem_a + stage * Cfg::A_BYTES, kb * Cfg::BLOCK_K,
                        wi.m_blk * kBlockM);
            tma_load_2d(&tmB, &full_bar[stage], smem_b + stage * Cfg::B_BYTES, kb * Cfg::BLOCK_K, nb * BN);
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (lane == 0) {
      constexpr uint32_t idesc = IS_FP8 ? umma_idesc(kBlockM, BN, 0, 0) : umma_idesc(kBlockM, BN, 1, 1);
      int stage = 0;
      uint32_t phase = 0;
      uint32_t tile_iter = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
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
            const uint64_t da = umma_desc_sw128_kmajor(smem_a + stage * Cfg::A_BYTES);
            const uint64_t db = umma_desc_sw128_kmajor(smem_b + stage * Cfg::B_BYTES);
#pragma unroll
            for (int k = 0; k < 4; ++k) {   // 4 x 32-byte K slices inside the 128-byte swizzle row
              const uint32_t acc = (kb | k) != 0;
              if constexpr (IS_FP8) umma_f8_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, acc);
              else umma_bf16_ss(tmem_d, da + 2 * k, db + 2 * k, idesc, acc);
            }
            umma_commit(&empty_bar[stage]);        // smem slot reusable once these MMAs retire
            if (++stage == kStages) { stage = 0; phase ^= 1; }
          }
          umma_commit(&tmem_full[as]);             // accumulator complete -> epilogue
        }
      }
    }
  } else if (warp >= 6) {
    // ================================ peer-copy warps =============================
    Gate::run(gp, static_cast<int>(threadIdx.x) - kNumThreads);
  } else {
    // ================================ epilogue ====================================
    const int quad = warp & 3;                       // TMEM lane quadrant this warp may read
    const int row_in_tile = quad * 32 + lane;
    Epi epi;
    uint32_t tile_iter = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      const WorkInfo wi = decode_work(w, num_m_blocks, num_n_blocks, n_splits, nb_per_split, sc);
      const int row = wi.m_blk * kBlockM + row_in_tile;
      epi.begin(ep, row, M, wi);
      for (int nb = wi.n_begin; nb < wi.n_end; ++nb, ++tile_iter) {
        const uint32_t as = tile_iter & 1;
        const uint32_t aphase = (tile_iter >> 1) & 1;
        mbar_wait(&tmem_full[as], aphase);
        tc_fence_after();
        AccReader acc{tmem_base + as * BN + (static_cast<uint32_t>(quad * 32) << 16)};
        epi.template tile<BN>(ep, acc, row, M, rot_n(nb, num_n_blocks, sc) * BN, N);
        tc_fence_before();
        mbar_arrive(&tmem_empty[as]);
      }
      epi.end(ep, row, M, wi);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------
// host helpers
// ------------------------------------------------------------------------------------------
using EncodeTiledFn = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                   const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                   CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D K-major operand: `rows` x `cols` elements, row pitch `ld_bytes`; box = 128 bytes x box_rows.
inline int make_tmap_kmajor(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_bytes,
                            uint32_t box_rows, int elem_bytes) {
  ensure_thread_ctx();
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    fprintf(stderr, "[prl] cuTensorMapEncodeTiled entry point not found\n");
    return -1;
  }
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_bytes};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(128 / elem_bytes), box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUtensorMapDataType dt = elem_bytes == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8;
  CUresult r = fn(m, dt, 2, const_cast<void*>(ptr), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    fprintf(stderr, "[prl] cuTensorMapEncodeTiled -> %d (ptr=%p rows=%llu cols=%llu ld_bytes=%llu box_rows=%u elem=%d)\n",
            static_cast<int>(r), ptr, static_cast<unsigned long long>(rows), static_cast<unsigned long long>(cols),
            static_cast<unsigned long long>(ld_bytes), box_rows, elem_bytes);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r);
}

inline int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
}

template <class Cfg, class Epi, class Gate, bool IS_FP8>
int launch_gemm(const void* A, const void* B, int M, int N, int K, int64_t lda_bytes, int64_t ldb_bytes,
                const typename Epi::Params& ep, const typename Gate::Params& gp, int n_splits, int max_ctas,
                cudaStream_t stream, Sched sc = Sched{0, 1, 0}, bool full_grid = false) {
  constexpr int EB = IS_FP8 ? 1 : 2;
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  CUtensorMap tmA, tmB;
  if (make_tmap_kmajor(&tmA, A, M, K, lda_bytes, kBlockM, EB)) return -2;
  if (make_tmap_kmajor(&tmB, B, N, K, ldb_bytes, Cfg::BLOCK_N, EB)) return -3;
  const int num_m_blocks = (M + kBlockM - 1) / kBlockM;
  const int num_n_blocks = (N + Cfg::BLOCK_N - 1) / Cfg::BLOCK_N;
  if (n_splits <= 0 || n_splits > num_n_blocks) n_splits = num_n_blocks;
  const int nb_per_split = (num_n_blocks + n_splits - 1) / n_splits;
  n_splits = (num_n_blocks + nb_per_split - 1) / nb_per_split;
  auto kern = gemm_sm100_kernel<Cfg, Epi, Gate, IS_FP8>;
  static bool attr_set = false;
  if (!attr_set) {
    PRL_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES));
    attr_set = true;
  }
  int ctas = sm_count();
  if (max_ctas > 0 && max_ctas < ctas) ctas = max_ctas;
  const int total = num_m_blocks * n_splits;
  if (total < ctas && !full_grid) ctas = total;
  kern<<<ctas, kNumThreads + 32 * Gate::kCommWarps, Cfg::SMEM_BYTES, stream>>>(
      tmA, tmB, ep, gp, M, N, K, num_m_blocks, num_n_blocks, n_splits, nb_per_split, sc);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

}  // namespace prl
