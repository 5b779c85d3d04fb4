// polyrl_b200 — ZeRO-3 parameter all-gather FUSED into the first consuming GEMM (one kernel).
//
//   D[M, W*R] = A[M,K] · concat_r(shard_r[R,K])^T        shard_r lives in rank r's HBM (NVLink peer)
//
// Every CTA carries two extra "peer-copy" warps next to the TMA / tcgen05 / epilogue roles.  The copy
// warps of all CTAs cooperatively pull the remote shards chunk by chunk (ld.global over NVLink ->
// st.global into the local gathered buffer), publishing one arrival counter per chunk
// (red.release.gpu).  The TMA producer of a tile whose weight rows live in chunk q acquires that
// counter before issuing its loads, and tiles are rasterised in arrival order (own shard first, then
// ring order), so the tcgen05 mainloop on shard r overlaps the NVLink transfer of shard r+1.  The
// gathered buffer is also what the backward pass re-uses (no second gather when it is kept).
//
// Baseline being replaced: torch FSDP all_gather_into_tensor + copy-out, then a cuBLAS GEMM
// (3rdparty/verl/verl/workers/fsdp_workers.py:494-507).
#include "gemm_sm100.cuh"

namespace prl {

constexpr int kAgMaxPeers = 16;

struct AgComm {
  static constexpr int kCommWarps = 2;
  struct Params {
    const int4* shard[kAgMaxPeers];   // peer shard base pointers (index = rank)
    int4* gathered;                   // local [W*R, K] buffer
    uint32_t* flags;                  // [W * chunks] arrival counters (monotone across launches)
    uint32_t target;                  // epoch * gridDim.x
    int64_t shard_n16;                // 16-byte units per shard
    int64_t chunk_n16;                // 16-byte units per chunk
    int chunks, world, rank;
    int nblocks_per_rank, nblocks_per_chunk;
    int local_in_place;               // shard[rank] already aliases gathered[rank]
  };

  __device__ __forceinline__ static void wait_n_block(const Params& p, int nb) {
    const int r = nb / p.nblocks_per_rank;
    int ring = r - p.rank;
    if (ring < 0) ring += p.world;
    if (ring == 0 && p.local_in_place) return;
    const int j = (nb - r * p.nblocks_per_rank) / p.nblocks_per_chunk;
    const uint32_t* f = p.flags + ring * p.chunks + j;
    uint32_t spins = 0;
    for (;;) {
      uint32_t v;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
      if (static_cast<int32_t>(v - p.target) >= 0) break;
      if (++spins > PRL_SPIN_LIMIT) { printf("[prl] ag_gemm chunk wait timeout\n"); __trap(); }
      __nanosleep(100);
    }
    asm volatile("fence.proxy.async.global;" ::: "memory");   // generic-proxy writes -> visible to TMA
  }

  __device__ __forceinline__ static void run(const Params& p, int tid) {
    constexpr int NT = kCommWarps * 32;
    const int64_t per_cta = (p.chunk_n16 + gridDim.x - 1) / gridDim.x;
    for (int ring = 0; ring < p.world; ++ring) {
      if (ring == 0 && p.local_in_place) continue;
      int r = p.rank + ring;
      if (r >= p.world) r -= p.world;
      const int4* src = p.shard[r];
      int4* dst = p.gathered + static_cast<int64_t>(r) * p.shard_n16;
      for (int j = 0; j < p.chunks; ++j) {
        const int64_t c0 = static_cast<int64_t>(j) * p.chunk_n16;
        const int64_t c_end = min(c0 + p.chunk_n16, p.shard_n16);
        const int64_t lo = c0 + static_cast<int64_t>(blockIdx.x) * per_cta;
        const int64_t hi = min(lo + per_cta, c_end);
        int64_t i = lo + tid;
        for (; i + 3 * NT < hi; i += 4 * NT) {
          const int4 a = ld_v4(src + i), b = ld_v4(src + i + NT), c = ld_v4(src + i + 2 * NT), d = ld_v4(src + i + 3 * NT);
          st_v4(dst + i, a); st_v4(dst + i + NT, b); st_v4(dst + i + 2 * NT, c); st_v4(dst + i + 3 * NT, d);
        }
        for (; i < hi; i += NT) st_v4(dst + i, ld_v4(src + i));
        __threadfence();
        asm volatile("bar.sync 1, %0;" ::"n"(NT) : "memory");
        if (tid == 0) {
          asm volatile("fence.proxy.async.global;" ::: "memory");
          asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p.flags + ring * p.chunks + j) : "memory");
        }
      }
    }
  }
};

// Same plain-store epilogue as gemm_tcgen05.cu (kept local so this TU is self-contained).
struct AgStoreParams {
  __nv_bfloat16* D;
  int64_t ldd;
};
struct EpiAgStore {
  using Params = AgStoreParams;
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
      __nv_bfloat16* dst = p.D + static_cast<int64_t>(row) * p.ldd + col;
      if (col + 32 <= N) {
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
        for (int i = 0; i < N - col; ++i) dst[i] = __float2bfloat16(v[i]);
      }
    }
  }
};

}  // namespace prl

using namespace prl;

// A:[M,K] bf16; shard_ptrs[r] -> rank r's [R,K] bf16 shard (peer-mapped); gathered: local [W*R,K];
// flags: uint32[W*chunks] zero-initialised once; epoch: 1,2,3,... per launch on this flags buffer.
extern "C" int prl_ag_gemm_bf16(const void* A, void* const* shard_ptrs, void* gathered, void* D, uint32_t* flags,
                                uint32_t epoch, int M, int N, int K, int64_t lda, int64_t ldd, int world, int rank,
                                int chunks, cudaStream_t stream) {
  using Cfg = GemmCfg<256, 2>;
  if (world > kAgMaxPeers || N % world != 0) return -50;
  const int R = N / world;
  if (R % Cfg::BLOCK_N != 0) return -51;              // n-blocks must not straddle ranks
  const int nblocks_per_rank = R / Cfg::BLOCK_N;
  if (chunks < 1) chunks = 1;
  while (nblocks_per_rank % chunks != 0) --chunks;    // chunks must tile the n-blocks of a shard
  AgComm::Params gp;
  for (int i = 0; i < kAgMaxPeers; ++i) gp.shard[i] = i < world ? static_cast<const int4*>(shard_ptrs[i]) : nullptr;
  gp.gathered = static_cast<int4*>(gathered);
  gp.flags = flags;
  const int ctas = sm_count();
  gp.target = epoch * static_cast<uint32_t>(ctas);
  gp.shard_n16 = static_cast<int64_t>(R) * K * 2 / 16;
  gp.chunk_n16 = gp.shard_n16 / chunks;
  gp.chunks = chunks;
  gp.world = world;
  gp.rank = rank;
  gp.nblocks_per_rank = nblocks_per_rank;
  gp.nblocks_per_chunk = nblocks_per_rank / chunks;
  gp.local_in_place =
      (static_cast<const char*>(shard_ptrs[rank]) == static_cast<const char*>(gathered) + static_cast<int64_t>(rank) * R * K * 2) ? 1 : 0;
  AgStoreParams ep{static_cast<__nv_bfloat16*>(D), ldd};
  Sched sc{1, gp.nblocks_per_chunk, rank * nblocks_per_rank};
  return launch_gemm<Cfg, EpiAgStore, AgComm, false>(A, gathered, M, N, K, lda * 2, static_cast<int64_t>(K) * 2, ep, gp,
                                                     0, 0, stream, sc, /*full_grid=*/true);
}
