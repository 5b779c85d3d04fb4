// polyrl_b200 — varlen causal flash-attention FORWARD on tcgen05 / TMEM / TMA (head_dim 128, GQA).
//
// Replaces the flash-attn 2 library call (mma.sync) the reference reaches through
// 3rdparty/verl/verl/models/transformers/monkey_patch.py:49-117 (`_flash_attention_forward`) for training,
// old-log-prob and prefill (incl. prefill behind a cached prefix: bottom-right aligned causal mask).
//
// One CTA = one (128-row query tile of one sequence, one query head), 192 threads:
//   warp 0   TMA producer: Q tile once, then K/V tiles (128 keys) through a 2-stage ring
//   warp 1   MMA issuer:   S_j = Q K_j^T (8 x tcgen05.mma 128x128x16, both operands K-major) into one of two TMEM
//                          S buffers — S_{j+1} is issued BEFORE waiting for P_j, so the tensor core works on the next
//                          scores while the softmax warps are busy — then  PV_j = P_j V_j  with P_j from shared
//                          memory and V_j consumed as an MN-major B operand (no transpose of V anywhere)
//   warps 2-9  softmax:    two warpgroups; group g owns columns [64g, 64g+64) of every query row (thread = TMEM lane):
//                          two passes over its 64 scores in TMEM (max — exchanged with the other group through shared
//                          memory and one 256-thread named barrier — then exp2 / sum), P_j written as bf16 into the
//                          128B-swizzled K-major layout the MMA reads, running output half kept in 64 fp32 registers:
//                          O = O * alpha + PV_j  (PV_j read back from TMEM).  exp2 is a single MUFU.EX2.
// Epilogue: O / l -> bf16 (16-byte stores), log-sum-exp per (head, token) for the backward pass.
#include "common.cuh"

namespace prl {

constexpr int kAfThreads = 320;          // warp0 TMA, warp1 MMA, warps 2-5 + 6-9: two softmax warpgroups
constexpr int kAfBM = 128, kAfBN = 128, kAfD = 128;
constexpr int kAfTile = 128 * 128 * 2;              // 32 KB: one 128 x 128 bf16 operand tile (2 chunks of 64 columns)
constexpr int kAfSmem = kAfTile /*Q*/ + 2 * 2 * kAfTile /*K,V x 2 stages*/ + kAfTile /*P*/ + 1024 + 256;

struct AttnFwdParams {
  __nv_bfloat16* o;            // [Tq, ldo]
  float* lse;                  // [nq, Tq]
  const int32_t* cu_q;         // [B+1]
  const int32_t* cu_k;         // [B+1]
  int64_t ldo;
  int B, nq, nkv, Tq;
  float scale_log2;            // sm_scale * log2(e)
  float sm_scale;
  int causal;
  long long* dbg;              // optional timeline buffer (kernel experiments)
};

__global__ void __launch_bounds__(kAfThreads, 1)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, AttnFwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem;
  uint8_t* sK = smem + kAfTile;                       // [2 stages]
  uint8_t* sV = smem + 3 * kAfTile;                   // [2 stages]
  uint8_t* sP = smem + 5 * kAfTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kAfTile);
  uint64_t* q_full = bars;            // 1
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* v_full = bars + 3;        // [2]
  uint64_t* kv_empty = bars + 5;      // [2]
  uint64_t* s_full = bars + 7;        // [2]
  uint64_t* s_empty = bars + 9;       // [2]  (128 arrivals)
  uint64_t* p_full = bars + 11;       // 128 arrivals
  uint64_t* o_full = bars + 12;       // 1
  uint64_t* o_empty = bars + 13;      // 128 arrivals
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 14);
  __shared__ int s_seq, s_q0;
  __shared__ float s_mx[2][2][128];             // [tile parity][warpgroup][row]: half-row maxima / sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y;
  const int hk = h / (p.nq / p.nkv);

  // ---- which (sequence, query tile) is this CTA?  tiles are numbered sequence by sequence
  if (threadIdx.x == 0) {
    int t = blockIdx.x, seq = -1, q0 = 0;
    for (int b = 0; b < p.B; ++b) {
      const int lq = p.cu_q[b + 1] - p.cu_q[b];
      const int nt = (lq + kAfBM - 1) / kAfBM;
      if (t < nt) { seq = b; q0 = t * kAfBM; break; }
      t -= nt;
    }
    s_seq = seq; s_q0 = q0;
  }
  __syncthreads();
  const int seq = s_seq;
  if (seq < 0) return;                                 // grid is an upper bound on the tile count
  const int q0 = s_q0;
  const int qs = p.cu_q[seq], lq = p.cu_q[seq + 1] - qs;
  const int ks = p.cu_k[seq], lk = p.cu_k[seq + 1] - ks;
  const int off = lk - lq;                             // bottom-right aligned causal mask (prefix-cached prefill: lk > lq)
  int kv_end = lk;
  if (p.causal) kv_end = min(lk, q0 + kAfBM + off);    // keys beyond the last row's diagonal are never needed
  const int nt = max(1, (kv_end + kAfBN - 1) / kAfBN);

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1); mbar_init(&v_full[i], 1); mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1); mbar_init(&s_empty[i], 256);
    }
    mbar_init(p_full, 256); mbar_init(o_full, 1); mbar_init(o_empty, 256);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS0 = tmem_base, tO = tmem_base + 256;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(q_full, kAfTile);
      tma_load_2d(&tmQ, q_full, sQ, h * kAfD, qs + q0);
      tma_load_2d(&tmQ, q_full, sQ + 16384, h * kAfD + 64, qs + q0);
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&kv_empty[s], ph ^ 1);
        mbar_expect_tx(&k_full[s], kAfTile);
        tma_load_2d(&tmK, &k_full[s], sK + s * kAfTile, hk * kAfD, ks + j * kAfBN);
        tma_load_2d(&tmK, &k_full[s], sK + s * kAfTile + 16384, hk * kAfD + 64, ks + j * kAfBN);
        mbar_expect_tx(&v_full[s], kAfTile);
        tma_load_2d(&tmV, &v_full[s], sV + s * kAfTile, hk * kAfD, ks + j * kAfBN);
        tma_load_2d(&tmV, &v_full[s], sV + s * kAfTile + 16384, hk * kAfD + 64, ks + j * kAfBN);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idS = umma_idesc(128, 128, 1, 1);
      constexpr uint32_t idO = umma_idesc_major(128, 128, 1, 1, 0, 1);      // B = V consumed MN-major
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&k_full[s], (j >> 1) & 1);
        mbar_wait(&s_empty[s], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t da = umma_desc_sw128_kmajor(sQ + (kk >> 2) * 16384) + 2 * (kk & 3);
          const uint64_t db = umma_desc_sw128_kmajor(sK + s * kAfTile + (kk >> 2) * 16384) + 2 * (kk & 3);
          umma_bf16_ss(tS0 + s * 128, da, db, idS, kk != 0);
        }
        umma_commit(&s_full[s]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < nt; ++j) {
        const int s = j & 1;
        if (j + 1 < nt) issue_s(j + 1);
        mbar_wait(p_full, j & 1);
        mbar_wait(&v_full[s], (j >> 1) & 1);
        mbar_wait(o_empty, (j & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {
          const uint64_t da = umma_desc_sw128_kmajor(sP + (kk >> 2) * 16384) + 2 * (kk & 3);
          const uint64_t db = umma_desc_sw128_mnmajor(sV + s * kAfTile, 16384) + kk * 128;
          umma_bf16_ss(tO, da, db, idO, kk != 0);
        }
        umma_commit(o_full);
        umma_commit(&kv_empty[s]);
      }
    }
  } else {
    // two softmax warpgroups: group g owns score / output columns [64g, 64g+64) of every row
    const int g = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // query row inside the tile == TMEM lane
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const int qrow = q0 + r;                              // position inside the sequence's queries
    const int kmax = p.causal ? qrow + off : lk - 1;      // last key this row may attend to
    float m_run = -INFINITY, l_run = 0.f;
    float O[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) O[i] = 0.f;
    const uint32_t sP_u32 = smem_u32(sP) + g * 16384 + r * 128;
    const bool rec = p.dbg != nullptr && blockIdx.x == gridDim.x / 2 && blockIdx.y == 0 && threadIdx.x == 64;
    long long* dbg = p.dbg;
    if (rec) dbg[0] = clock64();
    for (int j = 0; j < nt; ++j) {
      const int s = j & 1;
      mbar_wait(&s_full[s], (j >> 1) & 1);
      tc_fence_after();
      if (rec && j < 30) dbg[1 + 5 * j] = clock64();
      const uint32_t tS = tS0 + s * 128 + lane_base + g * 64;
      const int col0 = j * kAfBN + g * 64;
      const bool need_mask = (j * kAfBN + kAfBN - 1 > kmax) || (j * kAfBN + kAfBN > lk);
      // pass 1: max over this thread's 64 scores, then exchange with the other warpgroup's half of the row
      float mx = -INFINITY;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tS + c, v);
        tmem_ld_wait();
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = col0 + c + i;
            mx = fmaxf(mx, (col <= kmax && col < lk) ? __uint_as_float(v[i]) : -INFINITY);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, __uint_as_float(v[i]));
        }
      }
      s_mx[j & 1][g][r] = mx;
      asm volatile("bar.sync 1, 256;" ::: "memory");
      mx = fmaxf(mx, s_mx[j & 1][g ^ 1][r]);
      if (rec && j < 30) dbg[2 + 5 * j] = clock64();
      const float m_new = fmaxf(m_run, mx);
      const float m_use = (m_new == -INFINITY) ? 0.f : m_new;
      const float alpha = (m_run == -INFINITY) ? 0.f : fast_exp2((m_run - m_use) * p.scale_log2);
      const float mb = m_use * p.scale_log2;
      // pass 2: probabilities -> bf16 P in shared memory (K-major, 128B swizzle), partial row sum
      float lsum = 0.f;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tS + c, v);
        tmem_ld_wait();
        float pv[32];
        if (need_mask) {
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const int col = col0 + c + i;
            pv[i] = fast_exp2((col <= kmax && col < lk) ? __uint_as_float(v[i]) * p.scale_log2 - mb : -INFINITY);
          }
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) pv[i] = fast_exp2(__uint_as_float(v[i]) * p.scale_log2 - mb);
        }
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;          // four independent chains
#pragma unroll
        for (int i = 0; i < 32; i += 4) { s0 += pv[i]; s1 += pv[i + 1]; s2 += pv[i + 2]; s3 += pv[i + 3]; }
        lsum += (s0 + s1) + (s2 + s3);
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int u = (c >> 3) + q;                               // 16-byte unit inside this group's 128-byte row
          const uint32_t p0 = pack_bf16(pv[8 * q + 0], pv[8 * q + 1]), p1 = pack_bf16(pv[8 * q + 2], pv[8 * q + 3]);
          const uint32_t p2 = pack_bf16(pv[8 * q + 4], pv[8 * q + 5]), p3 = pack_bf16(pv[8 * q + 6], pv[8 * q + 7]);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP_u32 + ((u ^ (r & 7)) << 4)), "r"(p0), "r"(p1),
                       "r"(p2), "r"(p3)
                       : "memory");
        }
      }
      if (rec && j < 30) dbg[3 + 5 * j] = clock64();
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(p_full);
      mbar_arrive(&s_empty[s]);
      l_run = l_run * alpha + lsum;
      m_run = m_new;
      // O[:, 64g:64g+64] = O * alpha + (P V)[:, 64g:64g+64]
      mbar_wait(o_full, j & 1);
      tc_fence_after();
      if (rec && j < 30) dbg[4 + 5 * j] = clock64();
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tO + lane_base + g * 64 + c, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) O[c + i] = O[c + i] * alpha + __uint_as_float(v[i]);
      }
      tc_fence_before();
      mbar_arrive(o_empty);
      if (rec && j < 30) dbg[5 + 5 * j] = clock64();
    }
    // total row sum = the two groups' partial sums (same running max on both sides)
    s_mx[0][g][r] = l_run;
    asm volatile("bar.sync 1, 256;" ::: "memory");
    const float l_tot = l_run + s_mx[0][g ^ 1][r];
    if (qrow < lq) {
      const float inv = l_tot > 0.f ? 1.f / l_tot : 0.f;
      __nv_bfloat16* dst = p.o + static_cast<int64_t>(qs + qrow) * p.ldo + h * kAfD + g * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 8) {
        int4 ov;
        ov.x = pack_bf16(O[c] * inv, O[c + 1] * inv); ov.y = pack_bf16(O[c + 2] * inv, O[c + 3] * inv);
        ov.z = pack_bf16(O[c + 4] * inv, O[c + 5] * inv); ov.w = pack_bf16(O[c + 6] * inv, O[c + 7] * inv);
        st_v4(dst + c, ov);
      }
      if (p.lse && g == 0)
        p.lse[static_cast<int64_t>(h) * p.Tq + qs + qrow] = (m_run == -INFINITY) ? -INFINITY : m_run * p.sm_scale + __logf(l_tot);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

using EncodeTiledFn4 = CUresult (*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn4 af_encode_fn() {
  static EncodeTiledFn4 fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &qr) != cudaSuccess || !q) return nullptr;
    fn = reinterpret_cast<EncodeTiledFn4>(q);
  }
  return fn;
}
// rows x cols bf16 matrix with row pitch ld_bytes; box = 64 columns (128 B) x box_rows rows, 128B swizzle
int af_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_bytes, uint32_t box_rows) {
  ensure_thread_ctx();
  auto fn = af_encode_fn();
  if (!fn) return -1;
  cuuint64_t gdim[2] = {cols, rows};
  cuuint64_t gstride[1] = {ld_bytes};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -static_cast<int>(r);
}

}  // namespace prl

using namespace prl;

// q:[Tq, nq*128] (pitch ldq), k/v:[Tk, nkv*128] (pitches ldk/ldv), o:[Tq, nq*128] (pitch ldo); pitches in ELEMENTS.
// cu_q / cu_k: int32 [B+1] device arrays.  max_tiles: upper bound on sum_b ceil(lq_b / 128) (e.g. Tq/128 + B).
extern "C" int prl_attn_fwd(const void* q, const void* k, const void* v, void* o, float* lse, const int32_t* cu_q,
                            const int32_t* cu_k, int B, int Tq, int Tk, int nq, int nkv, int64_t ldq, int64_t ldk, int64_t ldv,
                            int64_t ldo, float sm_scale, int causal, int max_tiles, long long* dbg, cudaStream_t stream) {
  if (Tq <= 0 || B <= 0) return 0;
  if (nq % nkv != 0) return -110;
  CUtensorMap tmQ, tmK, tmV;
  if (af_tmap(&tmQ, q, Tq, static_cast<uint64_t>(nq) * kAfD, ldq * 2, kAfBM)) return -111;
  if (af_tmap(&tmK, k, Tk, static_cast<uint64_t>(nkv) * kAfD, ldk * 2, kAfBN)) return -112;
  if (af_tmap(&tmV, v, Tk, static_cast<uint64_t>(nkv) * kAfD, ldv * 2, kAfBN)) return -113;
  AttnFwdParams p;
  p.o = static_cast<__nv_bfloat16*>(o);
  p.lse = lse;
  p.cu_q = cu_q; p.cu_k = cu_k;
  p.ldo = ldo; p.B = B; p.nq = nq; p.nkv = nkv; p.Tq = Tq;
  p.sm_scale = sm_scale;
  p.scale_log2 = sm_scale * 1.4426950408889634f;
  p.causal = causal;
  p.dbg = dbg;
  static bool attr = false;
  if (!attr) {
    PRL_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAfSmem));
    attr = true;
  }
  dim3 grid(max_tiles, nq);
  attn_fwd_kernel<<<grid, kAfThreads, kAfSmem, stream>>>(tmQ, tmK, tmV, p);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
