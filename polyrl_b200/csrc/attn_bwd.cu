// polyrl_b200 — varlen causal flash-attention BACKWARD on tcgen05 / TMEM / TMA (head_dim 128, GQA).
//
// Replaces flash-attn 2's backward (reference: the autograd of `_flash_attention_forward`,
// 3rdparty/verl/verl/models/transformers/monkey_patch.py:49-117).
//
// One CTA = one (128-key tile of one sequence, one KV head).  K_j and V_j stay in shared memory; the CTA walks every
// query head of the GQA group and every 128-row query tile that can see these keys, and per (head, tile):
//     S  = Q_i K_j^T                 tcgen05  A = Q  (K-major)   B = K  (K-major)      -> TMEM
//     dP = dO_i V_j^T                tcgen05  A = dO (K-major)   B = V  (K-major)      -> TMEM
//     P  = exp2(S c - lse),  dS = P (dP - D) sm_scale           256 threads, thread = query row / half of the keys
//     dV += P^T  dO_i                tcgen05  A = P  (MN-major)  B = dO (MN-major)     accumulates in TMEM over ALL tiles
//     dK += dS^T Q_i                 tcgen05  A = dS (MN-major)  B = Q  (MN-major)     accumulates in TMEM over ALL tiles
//     dQ_i = dS K_j                  tcgen05  A = dS (K-major)   B = K  (MN-major)     -> TMEM -> red.global.add.v4.f32
// The bf16 P / dS tiles are written ONCE in the 128B-swizzled [query row][key] layout and read through K-major AND
// MN-major descriptors (the same bytes), Q / dO / K likewise — no operand is ever transposed.  dK / dV need no atomics
// (one CTA owns the key tile for the whole GQA group); dQ is accumulated in fp32 with vectorised reductions.
// TMEM: S/dQ 128 + dP 128 + dV 128 + dK 128 = 512 columns.
#include "common.cuh"

namespace prl {

constexpr int kAbThreads = 320;          // warp0 TMA, warp1 MMA, warps 2-9 compute (two warpgroups)
constexpr int kAbTile = 128 * 128 * 2;   // 32 KB
constexpr int kAbSmem = 6 * kAbTile + 1024 + 256;

struct AttnBwdParams {
  float* dq_acc;               // [Tq, nq*128] fp32, zero-initialised
  __nv_bfloat16* dk;           // [Tk, nkv*128]
  __nv_bfloat16* dv;           // [Tk, nkv*128]
  const float* lse;            // [nq, Tq]
  const float* dsum;           // [nq, Tq]   D = rowsum(dO * O)
  const int32_t* cu_q;
  const int32_t* cu_k;
  int B, nq, nkv, Tq;
  float scale_log2, sm_scale;
  int causal;
};

int af_tmap(CUtensorMap* m, const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld_bytes, uint32_t box_rows);

__device__ __forceinline__ void red_add_v4(float* p, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}

__global__ void __launch_bounds__(kAbThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmdO, AttnBwdParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sK = smem;
  uint8_t* sV = smem + kAbTile;
  uint8_t* sQ = smem + 2 * kAbTile;
  uint8_t* sdO = smem + 3 * kAbTile;
  uint8_t* sP = smem + 4 * kAbTile;
  uint8_t* sdS = smem + 5 * kAbTile;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kAbTile);
  uint64_t* kv_full = bars;           // 1
  uint64_t* qdo_full = bars + 1;      // 1
  uint64_t* qdo_empty = bars + 2;     // 1   (MMA commit: Q/dO/P/dS tiles consumed)
  uint64_t* sdp_full = bars + 3;      // 1   (MMA commit: S and dP in TMEM)
  uint64_t* pds_full = bars + 4;      // 256 (P and dS written to shared memory)
  uint64_t* dq_full = bars + 5;       // 1   (MMA commit: dQ tile in TMEM; dV/dK accumulators up to date)
  uint64_t* s_free = bars + 6;        // 256 (dQ drained: the S/dQ and dP columns may be overwritten)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 7);
  __shared__ int s_seq, s_k0;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int hk = blockIdx.y;
  const int G = p.nq / p.nkv;

  if (threadIdx.x == 0) {
    int t = blockIdx.x, seq = -1, k0 = 0;
    for (int b = 0; b < p.B; ++b) {
      const int lk = p.cu_k[b + 1] - p.cu_k[b];
      const int nt = (lk + 127) / 128;
      if (t < nt) { seq = b; k0 = t * 128; break; }
      t -= nt;
    }
    s_seq = seq; s_k0 = k0;
  }
  __syncthreads();
  const int seq = s_seq;
  if (seq < 0) return;
  const int k0 = s_k0;
  const int qs = p.cu_q[seq], lq = p.cu_q[seq + 1] - qs;
  const int ks = p.cu_k[seq], lk = p.cu_k[seq + 1] - ks;
  const int off = lk - lq;
  // first query tile whose LAST row can see key k0:  row + off >= k0
  int i0 = 0;
  if (p.causal) i0 = max(0, (k0 - off) / 128);
  const int nqt = (lq + 127) / 128;
  const int n_i = max(0, nqt - i0);
  const int iters = n_i * G;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV); tma_prefetch_desc(&tmdO);
    mbar_init(kv_full, 1); mbar_init(qdo_full, 1); mbar_init(qdo_empty, 1); mbar_init(sdp_full, 1);
    mbar_init(pds_full, 256); mbar_init(dq_full, 1); mbar_init(s_free, 256);
    fence_barrier_init();
  }
  if (warp == 1) { tmem_alloc(tmem_slot, 512); tmem_relinquish(); }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tS = tmem_base, tdP = tmem_base + 128, tdV = tmem_base + 256, tdK = tmem_base + 384;

  if (warp == 0) {
    if (lane == 0) {
      mbar_expect_tx(kv_full, 2 * kAbTile);
      tma_load_2d(&tmK, kv_full, sK, hk * 128, ks + k0);
      tma_load_2d(&tmK, kv_full, sK + 16384, hk * 128 + 64, ks + k0);
      tma_load_2d(&tmV, kv_full, sV, hk * 128, ks + k0);
      tma_load_2d(&tmV, kv_full, sV + 16384, hk * 128 + 64, ks + k0);
      for (int it = 0; it < iters; ++it) {
        const int h = hk * G + it / n_i, i = i0 + it % n_i;
        mbar_wait(qdo_empty, (it & 1) ^ 1);
        mbar_expect_tx(qdo_full, 2 * kAbTile);
        tma_load_2d(&tmQ, qdo_full, sQ, h * 128, qs + i * 128);
        tma_load_2d(&tmQ, qdo_full, sQ + 16384, h * 128 + 64, qs + i * 128);
        tma_load_2d(&tmdO, qdo_full, sdO, h * 128, qs + i * 128);
        tma_load_2d(&tmdO, qdo_full, sdO + 16384, h * 128 + 64, qs + i * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idKK = umma_idesc(128, 128, 1, 1);                       // A, B K-major
      constexpr uint32_t idMM = umma_idesc_major(128, 128, 1, 1, 1, 1);           // A, B MN-major
      constexpr uint32_t idKM = umma_idesc_major(128, 128, 1, 1, 0, 1);           // A K-major, B MN-major
      mbar_wait(kv_full, 0);
      for (int it = 0; it < iters; ++it) {
        mbar_wait(qdo_full, it & 1);
        mbar_wait(s_free, (it & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {      // S = Q K^T   (reduction over d)
          const uint64_t da = umma_desc_sw128_kmajor(sQ + (kk >> 2) * 16384) + 2 * (kk & 3);
          const uint64_t db = umma_desc_sw128_kmajor(sK + (kk >> 2) * 16384) + 2 * (kk & 3);
          umma_bf16_ss(tS, da, db, idKK, kk != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {      // dP = dO V^T (reduction over d)
          const uint64_t da = umma_desc_sw128_kmajor(sdO + (kk >> 2) * 16384) + 2 * (kk & 3);
          const uint64_t db = umma_desc_sw128_kmajor(sV + (kk >> 2) * 16384) + 2 * (kk & 3);
          umma_bf16_ss(tdP, da, db, idKK, kk != 0);
        }
        umma_commit(sdp_full);
        mbar_wait(pds_full, it & 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {      // dV += P^T dO  (reduction over the 128 query rows)
          const uint64_t da = umma_desc_sw128_mnmajor(sP, 16384) + kk * 128;
          const uint64_t db = umma_desc_sw128_mnmajor(sdO, 16384) + kk * 128;
          umma_bf16_ss(tdV, da, db, idMM, (it | kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {      // dK += dS^T Q
          const uint64_t da = umma_desc_sw128_mnmajor(sdS, 16384) + kk * 128;
          const uint64_t db = umma_desc_sw128_mnmajor(sQ, 16384) + kk * 128;
          umma_bf16_ss(tdK, da, db, idMM, (it | kk) != 0);
        }
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) {      // dQ = dS K    (reduction over the 128 keys)
          const uint64_t da = umma_desc_sw128_kmajor(sdS + (kk >> 2) * 16384) + 2 * (kk & 3);
          const uint64_t db = umma_desc_sw128_mnmajor(sK, 16384) + kk * 128;
          umma_bf16_ss(tS, da, db, idKM, kk != 0);
        }
        umma_commit(dq_full);
        umma_commit(qdo_empty);
      }
    }
  } else {
    const int g = (warp - 2) >> 2;
    const int quad = warp & 3;
    const int r = quad * 32 + lane;                       // query row inside the tile (S / dP / dQ lanes); key row for dK / dV
    const uint32_t lane_base = static_cast<uint32_t>(quad * 32) << 16;
    const uint32_t sP_u32 = smem_u32(sP) + g * 16384 + r * 128;
    const uint32_t sdS_u32 = smem_u32(sdS) + g * 16384 + r * 128;
    constexpr float kL2e = 1.4426950408889634f;
    for (int it = 0; it < iters; ++it) {
      const int h = hk * G + it / n_i, i = i0 + it % n_i;
      const int qrow = i * 128 + r;
      const bool row_ok = qrow < lq;
      const float lse2 = row_ok ? p.lse[static_cast<int64_t>(h) * p.Tq + qs + qrow] * kL2e : INFINITY;
      const float dsum = row_ok ? p.dsum[static_cast<int64_t>(h) * p.Tq + qs + qrow] : 0.f;
      const int kmax = p.causal ? qrow + off : lk - 1;
      mbar_wait(sdp_full, it & 1);
      tc_fence_after();
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t sv[32], dv[32];
        __syncwarp();
        tmem_ld_32x32(tS + lane_base + g * 64 + c, sv);
        tmem_ld_32x32(tdP + lane_base + g * 64 + c, dv);
        tmem_ld_wait();
        float pv[32], ds[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          const int col = k0 + g * 64 + c + e;
          const bool ok = col <= kmax && col < lk;
          const float pe = fast_exp2(ok ? __uint_as_float(sv[e]) * p.scale_log2 - lse2 : -INFINITY);
          pv[e] = pe;
          ds[e] = pe * (__uint_as_float(dv[e]) - dsum) * p.sm_scale;
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int u = (c >> 3) + q;
          const uint32_t sw = (u ^ (r & 7)) << 4;
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sP_u32 + sw), "r"(pack_bf16(pv[8 * q], pv[8 * q + 1])),
                       "r"(pack_bf16(pv[8 * q + 2], pv[8 * q + 3])), "r"(pack_bf16(pv[8 * q + 4], pv[8 * q + 5])),
                       "r"(pack_bf16(pv[8 * q + 6], pv[8 * q + 7]))
                       : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(sdS_u32 + sw), "r"(pack_bf16(ds[8 * q], ds[8 * q + 1])),
                       "r"(pack_bf16(ds[8 * q + 2], ds[8 * q + 3])), "r"(pack_bf16(ds[8 * q + 4], ds[8 * q + 5])),
                       "r"(pack_bf16(ds[8 * q + 6], ds[8 * q + 7]))
                       : "memory");
        }
      }
      fence_proxy_async();
      tc_fence_before();
      mbar_arrive(pds_full);
      // dQ tile -> fp32 accumulation buffer
      mbar_wait(dq_full, it & 1);
      tc_fence_after();
      float* dst = p.dq_acc + static_cast<int64_t>(qs + qrow) * (static_cast<int64_t>(p.nq) * 128) + h * 128 + g * 64;
#pragma unroll
      for (int c = 0; c < 64; c += 32) {
        uint32_t v[32];
        __syncwarp();
        tmem_ld_32x32(tS + lane_base + g * 64 + c, v);
        tmem_ld_wait();
        if (row_ok) {
#pragma unroll
          for (int e = 0; e < 32; e += 4)
            red_add_v4(dst + c + e, __uint_as_float(v[e]), __uint_as_float(v[e + 1]), __uint_as_float(v[e + 2]),
                       __uint_as_float(v[e + 3]));
        }
      }
      tc_fence_before();
      mbar_arrive(s_free);
    }
    // dK / dV of this key tile: rows = keys, this group's 64 columns
    if (iters > 0) {
      // the last dq_full wait above also covers the final dV / dK MMAs (same commit)
      tc_fence_after();
    }
    const int krow = k0 + r;
    __nv_bfloat16* dkp = p.dk + static_cast<int64_t>(ks + krow) * (static_cast<int64_t>(p.nkv) * 128) + hk * 128 + g * 64;
    __nv_bfloat16* dvp = p.dv + static_cast<int64_t>(ks + krow) * (static_cast<int64_t>(p.nkv) * 128) + hk * 128 + g * 64;
#pragma unroll
    for (int c = 0; c < 64; c += 32) {
      uint32_t a[32], b[32];
      if (iters > 0) {                       // warp-uniform: the TMEM loads are warp collectives, only the stores are per-row
        __syncwarp();
        tmem_ld_32x32(tdK + lane_base + g * 64 + c, a);
        tmem_ld_32x32(tdV + lane_base + g * 64 + c, b);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int e = 0; e < 32; ++e) { a[e] = 0; b[e] = 0; }
      }
      if (krow < lk) {
#pragma unroll
        for (int e = 0; e < 32; e += 8) {
          int4 o;
          o.x = pack_bf16(__uint_as_float(a[e]), __uint_as_float(a[e + 1])); o.y = pack_bf16(__uint_as_float(a[e + 2]), __uint_as_float(a[e + 3]));
          o.z = pack_bf16(__uint_as_float(a[e + 4]), __uint_as_float(a[e + 5])); o.w = pack_bf16(__uint_as_float(a[e + 6]), __uint_as_float(a[e + 7]));
          st_v4(dkp + c + e, o);
          o.x = pack_bf16(__uint_as_float(b[e]), __uint_as_float(b[e + 1])); o.y = pack_bf16(__uint_as_float(b[e + 2]), __uint_as_float(b[e + 3]));
          o.z = pack_bf16(__uint_as_float(b[e + 4]), __uint_as_float(b[e + 5])); o.w = pack_bf16(__uint_as_float(b[e + 6]), __uint_as_float(b[e + 7]));
          st_v4(dvp + c + e, o);
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) { tc_fence_after(); tmem_dealloc(tmem_base, 512); }
}

// D[h, t] = sum_d dO[t, h, d] * O[t, h, d]   (one warp per (token, head))
__global__ void attn_dsum_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ d_o, int64_t ldo, int64_t lddo,
                                 float* __restrict__ dsum, int Tq, int nq) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= Tq * nq) return;
  const int t = w / nq, h = w % nq;
  const uint2 a = *reinterpret_cast<const uint2*>(o + static_cast<int64_t>(t) * ldo + h * 128 + lane * 4);
  const uint2 b = *reinterpret_cast<const uint2*>(d_o + static_cast<int64_t>(t) * lddo + h * 128 + lane * 4);
  const float2 a0 = unpack_bf16(a.x), a1 = unpack_bf16(a.y), b0 = unpack_bf16(b.x), b1 = unpack_bf16(b.y);
  float s = a0.x * b0.x + a0.y * b0.y + a1.x * b1.x + a1.y * b1.y;
  s = warp_sum(s);
  if (lane == 0) dsum[static_cast<int64_t>(h) * Tq + t] = s;
}

}  // namespace prl

using namespace prl;

// dq_acc: fp32 [Tq, nq*128] zero-initialised by the caller; dk / dv: bf16 [Tk, nkv*128] (fully overwritten).
extern "C" int prl_attn_bwd(const void* q, const void* k, const void* v, const void* o, const void* d_o, const float* lse,
                            float* dsum, float* dq_acc, void* dk, void* dv, const int32_t* cu_q, const int32_t* cu_k, int B,
                            int Tq, int Tk, int nq, int nkv, int64_t ldq, int64_t ldk, int64_t ldv, int64_t ldo, int64_t lddo,
                            float sm_scale, int causal, int max_tiles, cudaStream_t stream) {
  if (Tq <= 0 || Tk <= 0 || B <= 0) return 0;
  if (nq % nkv != 0) return -120;
  {
    const int64_t warps = static_cast<int64_t>(Tq) * nq;
    const int threads = 256;
    const int64_t blocks = (warps * 32 + threads - 1) / threads;
    attn_dsum_kernel<<<static_cast<unsigned>(blocks), threads, 0, stream>>>(static_cast<const __nv_bfloat16*>(o),
                                                                           static_cast<const __nv_bfloat16*>(d_o), ldo, lddo, dsum, Tq, nq);
    PRL_CUDA_CHECK(cudaGetLastError());
  }
  CUtensorMap tmQ, tmK, tmV, tmdO;
  if (af_tmap(&tmQ, q, Tq, static_cast<uint64_t>(nq) * 128, ldq * 2, 128)) return -121;
  if (af_tmap(&tmK, k, Tk, static_cast<uint64_t>(nkv) * 128, ldk * 2, 128)) return -122;
  if (af_tmap(&tmV, v, Tk, static_cast<uint64_t>(nkv) * 128, ldv * 2, 128)) return -123;
  if (af_tmap(&tmdO, d_o, Tq, static_cast<uint64_t>(nq) * 128, lddo * 2, 128)) return -124;
  AttnBwdParams p;
  p.dq_acc = dq_acc; p.dk = static_cast<__nv_bfloat16*>(dk); p.dv = static_cast<__nv_bfloat16*>(dv);
  p.lse = lse; p.dsum = dsum; p.cu_q = cu_q; p.cu_k = cu_k;
  p.B = B; p.nq = nq; p.nkv = nkv; p.Tq = Tq;
  p.sm_scale = sm_scale; p.scale_log2 = sm_scale * 1.4426950408889634f; p.causal = causal;
  static bool attr = false;
  if (!attr) {
    PRL_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAbSmem));
    attr = true;
  }
  dim3 grid(max_tiles, nkv);
  attn_bwd_kernel<<<grid, kAbThreads, kAbSmem, stream>>>(tmQ, tmK, tmV, tmdO, p);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
