// polyrl_b200 — paged-KV flash-decode attention (one query token per sequence, GQA-aware, split-KV).
//
// Replaces the SGLang attention backend the reference depends on for rollout decode
// (rlboost/verl_stream/workers/rollout/sglang_rollout/sglang_rollout_remote.py:139 pins `fa3`,
// an sm_90a-only kernel family).  Decode attention is HBM-bound (every step streams the whole KV
// cache once), so the design goal is one pass over K and V at full bandwidth:
//   * cache layout [num_blocks, n_kv, 16, d]: a page of one kv-head is one contiguous 4 KB (d=128) run
//   * one CTA per (sequence, kv head, split); all G = nq/nkv query heads of the group share every
//     K/V byte (loaded ONCE, straight from HBM into mma.sync fragments — no shared-memory staging:
//     the k-index of an MMA is a free permutation, so each lane's 16-byte load is wired to the
//     k-slots it feeds; V fragments are built with two PRMTs per register pair)
//   * fp32 online softmax in registers, split results merged by a second tiny kernel.
#include <cstdlib>

#include "common.cuh"

namespace prl {

constexpr int kPage = 16;           // tokens per KV page
constexpr int kAttnWarps = 4;

__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

// partial outputs: o_part [num_seqs, nq, splits, D] fp32, ml_part [num_seqs, nq, splits, 2] (max2, sum)
// MINB = minimum resident CTAs per SM (register budget): 3 -> 168 regs, no spills; 4 -> 128 regs, ~76 B of spills but a third
// more warps in flight to hide HBM latency.  Both are compiled; POLYRL_DECODE_MINB picks (default from the microbench).
template <int D, int MINB>
__global__ void __launch_bounds__(kAttnWarps * 32, MINB)
paged_decode_kernel(const __nv_bfloat16* __restrict__ q, int64_t ldq,     // [num_seqs, nq*D] (row pitch ldq)
                    const __nv_bfloat16* __restrict__ k_cache, const __nv_bfloat16* __restrict__ v_cache,
                    const int32_t* __restrict__ block_table, int64_t ld_bt, const int32_t* __restrict__ seq_lens,
                    int nq, int nkv, int splits, float scale_log2, float* __restrict__ o_part,
                    float* __restrict__ ml_part, __nv_bfloat16* __restrict__ out_direct, int64_t ldo) {
  static_assert(D % 64 == 0, "head dim must be a multiple of 64");
  constexpr int KC = D / 32;     // 32-wide d chunks for QK (two MMAs each)
  constexpr int DB = D / 64;     // 64-wide d blocks for PV (eight n-tiles each)
  const int G = nq / nkv;        // query heads per kv head (<= 16)
  const int split = blockIdx.x % splits;
  const int kvh = (blockIdx.x / splits) % nkv;
  const int seq = blockIdx.x / (splits * nkv);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gq = lane >> 2, qd = lane & 3;    // fragment row / quad index

  const int len = seq_lens[seq];
  const int num_pages = (len + kPage - 1) / kPage;
  const int pages_per_split = (num_pages + splits - 1) / splits;
  const int p_begin = split * pages_per_split;
  const int p_end = min(num_pages, p_begin + pages_per_split);

  // ---- Q fragments (A operand), rows = heads of the group, k-slots permuted to match 16 B K loads
  uint32_t qf[KC][2][4];
  {
    const __nv_bfloat16* qbase = q + static_cast<int64_t>(seq) * ldq + static_cast<int64_t>(kvh) * G * D;
#pragma unroll
    for (int kc = 0; kc < KC; ++kc) {
      const int dcol = 32 * kc + 8 * qd;
      uint4 lo = make_uint4(0, 0, 0, 0), hi = make_uint4(0, 0, 0, 0);
      if (gq < G) lo = *reinterpret_cast<const uint4*>(qbase + gq * D + dcol);
      if (gq + 8 < G) hi = *reinterpret_cast<const uint4*>(qbase + (gq + 8) * D + dcol);
      qf[kc][0][0] = lo.x; qf[kc][0][1] = hi.x; qf[kc][0][2] = lo.y; qf[kc][0][3] = hi.y;
      qf[kc][1][0] = lo.z; qf[kc][1][1] = hi.z; qf[kc][1][2] = lo.w; qf[kc][1][3] = hi.w;
    }
  }

  float o[DB * 8][4];
#pragma unroll
  for (int i = 0; i < DB * 8; ++i) { o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f; }
  float m_lo = -INFINITY, m_hi = -INFINITY, l_lo = 0.f, l_hi = 0.f;   // rows gq and gq+8

  const int32_t* bt = block_table + static_cast<int64_t>(seq) * ld_bt;
  constexpr int kLinesPerPage = kPage * D * 2 / 128;          // 128-byte lines in one K (or V) page of one kv head
  int64_t blk_next = (p_begin + warp < p_end) ? bt[p_begin + warp] : 0;
  for (int p = p_begin + warp; p < p_end; p += kAttnWarps) {
    const int64_t blk = blk_next;
    const __nv_bfloat16* kp = k_cache + ((blk * nkv + kvh) * kPage) * D;
    const __nv_bfloat16* vp = v_cache + ((blk * nkv + kvh) * kPage) * D;
    // ---- software prefetch: look up this warp's NEXT page now (breaks the block-table -> KV dependent-load chain) and pull
    // its K and V lines into L2 while the current page is being consumed (zero register cost, doubles bytes in flight)
    if (p + kAttnWarps < p_end) {
      blk_next = bt[p + kAttnWarps];
      if (lane < kLinesPerPage) {
        const int64_t off = ((blk_next * nkv + kvh) * kPage) * D + lane * 64;
        asm volatile("prefetch.global.L2 [%0];" ::"l"(k_cache + off));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(v_cache + off));
      }
    }
    // ---- issue every load of this page up front (16 x LDG.128 in flight per lane)
    uint4 kr[2][KC];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
      for (int kc = 0; kc < KC; ++kc)
        kr[j][kc] = *reinterpret_cast<const uint4*>(kp + (8 * j + gq) * D + 32 * kc + 8 * qd);
    uint4 vr[4][DB];
    const int vt[4] = {2 * qd, 2 * qd + 1, 8 + 2 * qd, 9 + 2 * qd};
#pragma unroll
    for (int t = 0; t < 4; ++t)
#pragma unroll
      for (int db = 0; db < DB; ++db)
        vr[t][db] = *reinterpret_cast<const uint4*>(vp + vt[t] * D + 64 * db + 8 * gq);

    // ---- S = Q K^T for 16 tokens (two n-tiles of 8)
    float s[2][4];
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
      for (int kc = 0; kc < KC; ++kc) {
        mma_16816(s[j], qf[kc][0], kr[j][kc].x, kr[j][kc].y);
        mma_16816(s[j], qf[kc][1], kr[j][kc].z, kr[j][kc].w);
      }
    }
    // ---- mask + online softmax (base-2 domain)
    const int tok0 = p * kPage;
    float mx_lo = -INFINITY, mx_hi = -INFINITY;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const bool ok = tok0 + 8 * j + 2 * qd + e < len;
        s[j][e] = ok ? s[j][e] * scale_log2 : -INFINITY;
        s[j][2 + e] = ok ? s[j][2 + e] * scale_log2 : -INFINITY;
        mx_lo = fmaxf(mx_lo, s[j][e]);
        mx_hi = fmaxf(mx_hi, s[j][2 + e]);
      }
    }
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 1));
    mx_lo = fmaxf(mx_lo, __shfl_xor_sync(0xffffffffu, mx_lo, 2));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 1));
    mx_hi = fmaxf(mx_hi, __shfl_xor_sync(0xffffffffu, mx_hi, 2));
    const float nm_lo = fmaxf(m_lo, mx_lo), nm_hi = fmaxf(m_hi, mx_hi);   // finite: page has >=1 valid token
    const float r_lo = exp2f(m_lo - nm_lo), r_hi = exp2f(m_hi - nm_hi);
    m_lo = nm_lo; m_hi = nm_hi;
    float ps_lo = 0.f, ps_hi = 0.f;
    uint32_t pf[4];
    {
      float e00 = exp2f(s[0][0] - nm_lo), e01 = exp2f(s[0][1] - nm_lo);
      float e02 = exp2f(s[0][2] - nm_hi), e03 = exp2f(s[0][3] - nm_hi);
      float e10 = exp2f(s[1][0] - nm_lo), e11 = exp2f(s[1][1] - nm_lo);
      float e12 = exp2f(s[1][2] - nm_hi), e13 = exp2f(s[1][3] - nm_hi);
      ps_lo = e00 + e01 + e10 + e11;
      ps_hi = e02 + e03 + e12 + e13;
      pf[0] = pack_bf16(e00, e01);   // row gq,   tokens 2qd,2qd+1
      pf[1] = pack_bf16(e02, e03);   // row gq+8
      pf[2] = pack_bf16(e10, e11);   // row gq,   tokens 8+2qd,9+2qd
      pf[3] = pack_bf16(e12, e13);
    }
    l_lo = l_lo * r_lo + ps_lo;      // per-lane partial row sums (reduced across the quad at the end)
    l_hi = l_hi * r_hi + ps_hi;
    // ---- O = O*r + P V
#pragma unroll
    for (int db = 0; db < DB; ++db) {
      const uint32_t* v0 = reinterpret_cast<const uint32_t*>(&vr[0][db]);
      const uint32_t* v1 = reinterpret_cast<const uint32_t*>(&vr[1][db]);
      const uint32_t* v2 = reinterpret_cast<const uint32_t*>(&vr[2][db]);
      const uint32_t* v3 = reinterpret_cast<const uint32_t*>(&vr[3][db]);
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const uint32_t sel = e ? 0x7632u : 0x5410u;
          const uint32_t b0 = prmt(v0[r], v1[r], sel);
          const uint32_t b1 = prmt(v2[r], v3[r], sel);
          float(&acc)[4] = o[db * 8 + 2 * r + e];
          acc[0] *= r_lo; acc[1] *= r_lo; acc[2] *= r_hi; acc[3] *= r_hi;
          mma_16816(acc, pf, b0, b1);
        }
      }
    }
  }

  // ---- reduce row sums across the quad
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 1);
  l_lo += __shfl_xor_sync(0xffffffffu, l_lo, 2);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 1);
  l_hi += __shfl_xor_sync(0xffffffffu, l_hi, 2);

  // ---- combine the 4 warps through shared memory
  __shared__ float sm_m[kAttnWarps][16], sm_l[kAttnWarps][16];
  __shared__ float sm_o[kAttnWarps][16][D + 4];
  if (qd == 0) {
    sm_m[warp][gq] = m_lo; sm_m[warp][gq + 8] = m_hi;
    sm_l[warp][gq] = l_lo; sm_l[warp][gq + 8] = l_hi;
  }
#pragma unroll
  for (int db = 0; db < DB; ++db)
#pragma unroll
    for (int nt = 0; nt < 8; ++nt) {
      const float(&acc)[4] = o[db * 8 + nt];
      // C layout: row gq (+8), cols 2qd,2qd+1 of the n-tile; actual d = 64db + 8*col + nt
      sm_o[warp][gq][64 * db + 8 * (2 * qd) + nt] = acc[0];
      sm_o[warp][gq][64 * db + 8 * (2 * qd + 1) + nt] = acc[1];
      sm_o[warp][gq + 8][64 * db + 8 * (2 * qd) + nt] = acc[2];
      sm_o[warp][gq + 8][64 * db + 8 * (2 * qd + 1) + nt] = acc[3];
    }
  __syncthreads();
  for (int idx = threadIdx.x; idx < G * D; idx += kAttnWarps * 32) {
    const int g = idx / D, dcol = idx - g * D;
    float mm = -INFINITY;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) mm = fmaxf(mm, sm_m[w][g]);
    float acc = 0.f, ll = 0.f;
#pragma unroll
    for (int w = 0; w < kAttnWarps; ++w) {
      const float f = (sm_m[w][g] == -INFINITY) ? 0.f : exp2f(sm_m[w][g] - mm);
      acc += sm_o[w][g][dcol] * f;
      ll += sm_l[w][g] * f;
    }
    const int head = kvh * G + g;
    if (out_direct) {        // single split: normalise and write the final bf16 output here (no merge kernel)
      out_direct[static_cast<int64_t>(seq) * ldo + head * D + dcol] = __float2bfloat16(ll > 0.f ? acc / ll : 0.f);
      continue;
    }
    const int64_t base = (static_cast<int64_t>(seq) * nq + head) * splits + split;
    o_part[base * D + dcol] = acc;
    if (dcol == 0) { ml_part[base * 2] = mm; ml_part[base * 2 + 1] = ll; }
  }
}

// merge splits -> bf16 output [num_seqs, nq*D]
template <int D>
__global__ void decode_merge_kernel(const float* __restrict__ o_part, const float* __restrict__ ml_part, int splits,
                                    __nv_bfloat16* __restrict__ out, int64_t ldo, int nq) {
  const int seq = blockIdx.x / nq, head = blockIdx.x % nq;
  const int64_t base = (static_cast<int64_t>(seq) * nq + head) * splits;
  float mm = -INFINITY;
  for (int s = 0; s < splits; ++s) mm = fmaxf(mm, ml_part[(base + s) * 2]);
  float ll = 0.f;
  for (int s = 0; s < splits; ++s) {
    const float m = ml_part[(base + s) * 2];
    ll += (m == -INFINITY) ? 0.f : ml_part[(base + s) * 2 + 1] * exp2f(m - mm);
  }
  const float inv = ll > 0.f ? 1.f / ll : 0.f;
  for (int dcol = threadIdx.x; dcol < D; dcol += blockDim.x) {
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) {
      const float m = ml_part[(base + s) * 2];
      if (m != -INFINITY) acc += o_part[(base + s) * D + dcol] * exp2f(m - mm);
    }
    out[static_cast<int64_t>(seq) * ldo + head * D + dcol] = __float2bfloat16(acc * inv);
  }
}

}  // namespace prl

using namespace prl;

extern "C" int prl_paged_decode_workspace_floats(int num_seqs, int nq, int d, int splits) {
  return num_seqs * nq * splits * (d + 2);
}

// q:[num_seqs, nq*d] bf16; k/v cache [num_blocks, nkv, 16, d] bf16; block_table int32 [num_seqs, ld_bt];
// seq_lens int32 (INCLUDING the current token, already appended to the cache).
extern "C" int prl_paged_decode(const void* q, int64_t ldq, const void* k_cache, const void* v_cache,
                                const int32_t* block_table, int64_t ld_bt, const int32_t* seq_lens, int num_seqs,
                                int nq, int nkv, int d, float sm_scale, int splits, float* workspace, void* out,
                                int64_t ldo, cudaStream_t s) {
  if (num_seqs <= 0) return 0;
  if (nq % nkv != 0 || nq / nkv > 16) return -30;
  if (splits < 1) splits = 1;
  float* o_part = workspace;
  float* ml_part = workspace + static_cast<int64_t>(num_seqs) * nq * splits * d;
  const float scale_log2 = sm_scale * 1.4426950408889634f;
  const int grid = num_seqs * nkv * splits;
  static const int minb = [] { const char* e = getenv("POLYRL_DECODE_MINB"); return e ? atoi(e) : 3; }();
  if (d == 128) {
    auto kern = minb >= 4 ? paged_decode_kernel<128, 4> : paged_decode_kernel<128, 3>;
    kern<<<grid, kAttnWarps * 32, 0, s>>>(
        static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k_cache),
        static_cast<const __nv_bfloat16*>(v_cache), block_table, ld_bt, seq_lens, nq, nkv, splits, scale_log2, o_part,
        ml_part, splits == 1 ? static_cast<__nv_bfloat16*>(out) : nullptr, ldo);
    if (splits > 1)
      decode_merge_kernel<128><<<num_seqs * nq, 128, 0, s>>>(o_part, ml_part, splits, static_cast<__nv_bfloat16*>(out), ldo, nq);
  } else if (d == 64) {
    paged_decode_kernel<64, 3><<<grid, kAttnWarps * 32, 0, s>>>(
        static_cast<const __nv_bfloat16*>(q), ldq, static_cast<const __nv_bfloat16*>(k_cache),
        static_cast<const __nv_bfloat16*>(v_cache), block_table, ld_bt, seq_lens, nq, nkv, splits, scale_log2, o_part,
        ml_part, splits == 1 ? static_cast<__nv_bfloat16*>(out) : nullptr, ldo);
    if (splits > 1)
      decode_merge_kernel<64><<<num_seqs * nq, 64, 0, s>>>(o_part, ml_part, splits, static_cast<__nv_bfloat16*>(out), ldo, nq);
  } else {
    return -31;
  }
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
