// polyrl_b200 — Python bindings (pybind11 through torch) for the sm_100a kernels in csrc/*.cu.
// Kernel translation units are torch-free (raw pointers + cudaStream_t) so they compile in seconds;
// this file is the only one that sees torch headers.
#include <ATen/cuda/CUDAContext.h>
#include "kv_manager.hpp"
#include "scheduler.hpp"
#include <c10/cuda/CUDAGuard.h>
#include <torch/extension.h>

#include <cstdint>
#include <string>
#include <vector>

extern "C" {
int prl_gemm_bf16(const void*, const void*, void*, const void*, int, int, int, int64_t, int64_t, int64_t, int, cudaStream_t);
int prl_gemm2_bf16(const void*, const void*, void*, const void*, int, int, int, int64_t, int64_t, int64_t, int, int, int, cudaStream_t);
int prl_gemm2_bf16_accum_f32(const void*, const void*, float*, float, int, int, int, int64_t, int64_t, int64_t, int, int, cudaStream_t);
int prl_gemm_fp8(const void*, const void*, void*, const float*, const float*, int, int, int, int64_t, int64_t, int64_t, int, cudaStream_t);
int prl_gemm_swiglu_bf16(const void*, const void*, void*, void*, int, int, int, int64_t, int64_t, int64_t, int64_t, cudaStream_t);
int prl_linear_ce_num_splits(int, int);
int prl_linear_ce_fwd(const void*, const void*, const int64_t*, float, int, int, int, int64_t, int64_t, float*, int, float*, float*, float*, float*, cudaStream_t);
int prl_linear_ce_dz(const void*, const void*, const int64_t*, const float*, const float*, const float*, const float*, float, int, int, int, int, int64_t, int64_t, void*, int64_t, cudaStream_t);
int prl_rmsnorm_fwd(const void*, const void*, void*, float*, int, int, float, cudaStream_t);
int prl_add_rmsnorm(void*, const void*, const void*, void*, int, int, float, cudaStream_t);
int prl_rmsnorm_bwd(const void*, const void*, const void*, const float*, void*, float*, int, int, cudaStream_t);
int prl_rope(void*, const int64_t*, const float*, const float*, int, int, int, int, int64_t, float, void*, void*, const int64_t*, int, cudaStream_t);
int prl_swiglu_bwd(const void*, const void*, void*, int64_t, int, cudaStream_t);
int prl_quant_fp8_rowwise(const void*, void*, float*, int, int, cudaStream_t);
int prl_grad_accum(float*, const void*, int, float, int64_t, float*, cudaStream_t);
int prl_sumsq(const float*, int64_t, float*, cudaStream_t);
int prl_adamw(float*, float*, float*, float*, void*, int64_t, float, float, float, float, float, int, float, const float*, int*, int, cudaStream_t);
int prl_paged_decode_workspace_floats(int, int, int, int);
int prl_paged_decode(const void*, int64_t, const void*, const void*, const int32_t*, int64_t, const int32_t*, int, int, int, int, float, int, float*, void*, int64_t, cudaStream_t);
int prl_sample(const void*, int, int64_t, int, int, const float*, const uint64_t*, int, int64_t*, float*, cudaStream_t);
int prl_sample_topk_topp(const void*, int, int64_t, int, int, const float*, const uint64_t*, int, int, float, int64_t*, float*, cudaStream_t);
int prl_attn_fwd(const void*, const void*, const void*, void*, float*, const int32_t*, const int32_t*, int, int, int, int, int, int64_t, int64_t, int64_t, int64_t, float, int, int, long long*, cudaStream_t);
int prl_attn_bwd(const void*, const void*, const void*, const void*, const void*, const float*, float*, float*, void*, void*, const int32_t*, const int32_t*, int, int, int, int, int, int64_t, int64_t, int64_t, int64_t, int64_t, float, int, int, cudaStream_t);
int prl_ipc_get_handle(void*, void*);
int prl_ipc_open_handle(const void*, void**);
int prl_ipc_close_handle(void*);
int prl_cuda_malloc(void**, int64_t);
int prl_cuda_free(void*);
int prl_enable_peer_access(int);
int prl_p2p_barrier(void* const*, int, int, uint32_t*, cudaStream_t);
int prl_p2p_copy(const void*, void*, int64_t, int, cudaStream_t);
int prl_p2p_push_fp8(const void*, void*, float*, int, int, int, cudaStream_t);
int prl_p2p_allgather_pull(void* const*, void*, int64_t, int, int, int, cudaStream_t);
int prl_p2p_allreduce(void* const*, int64_t, void*, const void*, int64_t, int, int, cudaStream_t);
int64_t prl_vmm_granularity(int);
int prl_vmm_reserve(int64_t, int, int64_t*, int64_t*);
int prl_vmm_map(int64_t);
int prl_vmm_unmap(int64_t);
int prl_vmm_free(int64_t);
int prl_vmm_is_mapped(int64_t);
int prl_p2p_reduce_scatter_accum(void* const*, int64_t, float*, int64_t, float, int, int, float*, int, cudaStream_t);
int prl_nvls_supported(int);
int64_t prl_nvls_round_size(int64_t, int, int);
int64_t prl_nvls_create(int64_t, int);
int64_t prl_nvls_mc_create(int64_t, int);
int prl_nvls_export_fd(int64_t);
int64_t prl_nvls_import_fd(int, int64_t, int);
int prl_nvls_mc_add_device(int64_t, int);
int prl_nvls_mc_bind(int64_t, int64_t);
int64_t prl_nvls_map(int64_t, int);
int prl_nvls_unmap(int64_t);
int prl_nvls_release(int64_t);
int prl_close_fd(int);
int prl_nvls_allreduce(const void*, void*, const void*, int64_t, void* const*, int, int, uint32_t*, int, cudaStream_t);
int prl_nvls_allreduce_oneshot(const void*, void*, const void*, int64_t, void* const*, int, int, uint32_t*, int, cudaStream_t);
int prl_nvls_reduce_scatter_accum(const void*, int64_t, float*, int64_t, float, float*, int, cudaStream_t);
int prl_nvls_bcast_store(const void*, void*, int64_t, int, cudaStream_t);
int prl_skinny_workspace_floats(int, int, int);
int prl_skinny_gemm(const void*, const void*, void*, const void*, const void*, const float*, const float*, float*, int*, int, int, int, int64_t, int64_t, int64_t, int, int, int, cudaStream_t);
int prl_stream_gemm(const void*, const void*, void*, const void*, const void*, const float*, const float*, int, int, int, int64_t, int64_t, int64_t, int, int, int, int, cudaStream_t);
int prl_skinny_gemm_tp(const void*, const void*, void*, const void*, const void*, const float*, const float*, float*, int*, int, int, int, int64_t, int64_t, int64_t, int, int, int, int, int, uint32_t*, int*, int64_t, void* const*, void* const*, cudaStream_t);
int prl_ag_gemm_bf16(const void*, void* const*, void*, void*, uint32_t*, uint32_t, int, int, int, int64_t, int64_t, int, int, int, cudaStream_t);
}

namespace {

inline cudaStream_t cur_stream() { return at::cuda::getCurrentCUDAStream().stream(); }
std::vector<void*> to_ptrs(const std::vector<int64_t>& v);

inline void check(int rc, const char* what) {
  TORCH_CHECK(rc == 0, "polyrl_b200 kernel '", what, "' failed with code ", rc);
}
#define CHECK_BF16(t) TORCH_CHECK((t).is_cuda() && (t).scalar_type() == at::kBFloat16, #t " must be a CUDA bf16 tensor")
#define CHECK_F32(t) TORCH_CHECK((t).is_cuda() && (t).scalar_type() == at::kFloat, #t " must be a CUDA fp32 tensor")
#define CHECK_LASTDIM(t) TORCH_CHECK((t).stride(-1) == 1, #t " must be contiguous in the last dim")

at::Tensor gemm_bf16(const at::Tensor& a, const at::Tensor& b, const c10::optional<at::Tensor>& residual,
                     c10::optional<at::Tensor> out, int64_t force_bn) {
  CHECK_BF16(a); CHECK_BF16(b);
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2 && a.size(1) == b.size(1), "gemm_bf16: A[M,K] B[N,K]");
  CHECK_LASTDIM(a); CHECK_LASTDIM(b);
  const int M = a.size(0), K = a.size(1), N = b.size(0);
  TORCH_CHECK(K % 8 == 0 && N % 8 == 0, "K and N must be multiples of 8");
  TORCH_CHECK(a.stride(0) % 8 == 0 && b.stride(0) % 8 == 0, "row pitch must be a multiple of 8 elements");
  c10::cuda::CUDAGuard g(a.device());
  at::Tensor d = out.has_value() ? *out : at::empty({M, N}, a.options());
  CHECK_BF16(d); CHECK_LASTDIM(d);
  const void* r = nullptr;
  if (residual.has_value()) {
    CHECK_BF16(*residual);
    TORCH_CHECK(residual->sizes() == d.sizes() && residual->stride(0) == d.stride(0) && residual->stride(1) == 1);
    r = residual->data_ptr();
  }
  check(prl_gemm_bf16(a.data_ptr(), b.data_ptr(), d.data_ptr(), r, M, N, K, a.stride(0), b.stride(0), d.stride(0),
                      static_cast<int>(force_bn), cur_stream()), "gemm_bf16");
  return d;
}

// CTA-pair (cta_group::2) GEMM family.  D[M,N] = sum_k A(m,k) B(n,k) (+residual).
//   a_mn=false: a is [M,K];  a_mn=true: a is [K,M] (consumed MN-major, never transposed in memory)
//   b_mn=false: b is [N,K];  b_mn=true: b is [K,N]
at::Tensor gemm2_bf16(const at::Tensor& a, const at::Tensor& b, bool a_mn, bool b_mn,
                      const c10::optional<at::Tensor>& residual, c10::optional<at::Tensor> out, int64_t force_bn) {
  CHECK_BF16(a); CHECK_BF16(b);
  TORCH_CHECK(a.dim() == 2 && b.dim() == 2, "gemm2_bf16: 2-D operands");
  CHECK_LASTDIM(a); CHECK_LASTDIM(b);
  const int M = a_mn ? a.size(1) : a.size(0), K = a_mn ? a.size(0) : a.size(1);
  const int N = b_mn ? b.size(1) : b.size(0), Kb = b_mn ? b.size(0) : b.size(1);
  TORCH_CHECK(K == Kb, "gemm2_bf16: reduction lengths differ");
  TORCH_CHECK(a.size(1) % 8 == 0 && b.size(1) % 8 == 0 && N % 8 == 0, "inner dims must be multiples of 8");
  TORCH_CHECK(a.stride(0) % 8 == 0 && b.stride(0) % 8 == 0, "row pitch must be a multiple of 8 elements");
  TORCH_CHECK((reinterpret_cast<uintptr_t>(a.data_ptr()) & 15) == 0 && (reinterpret_cast<uintptr_t>(b.data_ptr()) & 15) == 0,
              "operands must be 16-byte aligned");
  c10::cuda::CUDAGuard g(a.device());
  at::Tensor d = out.has_value() ? *out : at::empty({M, N}, a.options());
  CHECK_BF16(d); CHECK_LASTDIM(d);
  TORCH_CHECK(d.size(0) == M && d.size(1) == N && d.stride(0) % 8 == 0 && (reinterpret_cast<uintptr_t>(d.data_ptr()) & 15) == 0);
  const void* r = nullptr;
  if (residual.has_value()) {
    CHECK_BF16(*residual);
    TORCH_CHECK(residual->sizes() == d.sizes() && residual->stride(0) == d.stride(0) && residual->stride(1) == 1);
    r = residual->data_ptr();
  }
  TORCH_CHECK(a.stride(0) >= a.size(1) && b.stride(0) >= b.size(1), "gemm2_bf16: row pitch shorter than the row");
  const int rc = prl_gemm2_bf16(a.data_ptr(), b.data_ptr(), d.data_ptr(), r, M, N, K, a.stride(0), b.stride(0), d.stride(0),
                                a_mn ? 1 : 0, b_mn ? 1 : 0, static_cast<int>(force_bn), cur_stream());
  TORCH_CHECK(rc == 0, "polyrl_b200 kernel 'gemm2_bf16' failed with code ", rc, " (M=", M, " N=", N, " K=", K, " a_mn=", a_mn,
              " b_mn=", b_mn, " lda=", a.stride(0), " ldb=", b.stride(0), " ldd=", d.stride(0), " a=", a.data_ptr(), " b=",
              b.data_ptr(), " dev=", a.get_device(), ")");
  return d;
}

// out32 = beta * out32 + sum_k A(m,k) B(n,k)   (fp32 output, in place)
void gemm2_accum_f32_(at::Tensor out32, const at::Tensor& a, const at::Tensor& b, bool a_mn, bool b_mn, double beta) {
  CHECK_BF16(a); CHECK_BF16(b); CHECK_F32(out32); CHECK_LASTDIM(a); CHECK_LASTDIM(b); CHECK_LASTDIM(out32);
  const int M = a_mn ? a.size(1) : a.size(0), K = a_mn ? a.size(0) : a.size(1);
  const int N = b_mn ? b.size(1) : b.size(0), Kb = b_mn ? b.size(0) : b.size(1);
  TORCH_CHECK(K == Kb && out32.dim() == 2 && out32.size(0) == M && out32.size(1) == N);
  TORCH_CHECK(a.size(1) % 8 == 0 && b.size(1) % 8 == 0 && a.stride(0) % 8 == 0 && b.stride(0) % 8 == 0 && out32.stride(0) % 4 == 0);
  c10::cuda::CUDAGuard g(a.device());
  check(prl_gemm2_bf16_accum_f32(a.data_ptr(), b.data_ptr(), out32.data_ptr<float>(), static_cast<float>(beta), M, N, K,
                                 a.stride(0), b.stride(0), out32.stride(0), a_mn ? 1 : 0, b_mn ? 1 : 0, cur_stream()),
        "gemm2_accum_f32");
}

at::Tensor gemm_fp8(const at::Tensor& a, const at::Tensor& b, const at::Tensor& row_scale, const at::Tensor& col_scale,
                    int64_t force_bn) {
  TORCH_CHECK(a.is_cuda() && b.is_cuda() && a.element_size() == 1 && b.element_size() == 1, "fp8 operands expected");
  CHECK_F32(row_scale); CHECK_F32(col_scale);
  const int M = a.size(0), K = a.size(1), N = b.size(0);
  TORCH_CHECK(b.size(1) == K && K % 16 == 0 && N % 8 == 0);
  c10::cuda::CUDAGuard g(a.device());
  at::Tensor d = at::empty({M, N}, a.options().dtype(at::kBFloat16));
  check(prl_gemm_fp8(a.data_ptr(), b.data_ptr(), d.data_ptr(), row_scale.data_ptr<float>(),
                     col_scale.data_ptr<float>(), M, N, K, a.stride(0), b.stride(0), d.stride(0),
                     static_cast<int>(force_bn), cur_stream()), "gemm_fp8");
  return d;
}

std::vector<at::Tensor> gemm_swiglu(const at::Tensor& a, const at::Tensor& wgu, bool save_gu) {
  CHECK_BF16(a); CHECK_BF16(wgu); CHECK_LASTDIM(a); CHECK_LASTDIM(wgu);
  const int M = a.size(0), K = a.size(1), N2 = wgu.size(0);
  TORCH_CHECK(wgu.size(1) == K && N2 % 256 == 0 && K % 8 == 0);
  c10::cuda::CUDAGuard g(a.device());
  at::Tensor h = at::empty({M, N2 / 2}, a.options());
  at::Tensor gu = save_gu ? at::empty({M, N2}, a.options()) : at::Tensor();
  check(prl_gemm_swiglu_bf16(a.data_ptr(), wgu.data_ptr(), h.data_ptr(), save_gu ? gu.data_ptr() : nullptr, M, N2, K,
                             a.stride(0), wgu.stride(0), h.stride(0), save_gu ? gu.stride(0) : 0, cur_stream()),
        "gemm_swiglu");
  if (save_gu) return {h, gu};
  return {h};
}

std::vector<at::Tensor> linear_ce_fwd(const at::Tensor& hidden, const at::Tensor& w, const at::Tensor& labels,
                                      double inv_tau) {
  CHECK_BF16(hidden); CHECK_BF16(w); CHECK_LASTDIM(hidden); CHECK_LASTDIM(w);
  TORCH_CHECK(labels.is_cuda() && labels.scalar_type() == at::kLong && labels.is_contiguous());
  const int T = hidden.size(0), H = hidden.size(1), V = w.size(0);
  TORCH_CHECK(w.size(1) == H && labels.numel() == T && H % 8 == 0);
  c10::cuda::CUDAGuard g(hidden.device());
  const int S = prl_linear_ce_num_splits(T, V);
  auto fopt = hidden.options().dtype(at::kFloat);
  at::Tensor scratch = at::empty({static_cast<int64_t>(T) * (3 * S + 1)}, fopt);
  at::Tensor lp = at::empty({T}, fopt), ent = at::empty({T}, fopt), lse = at::empty({T}, fopt), ez = at::empty({T}, fopt);
  check(prl_linear_ce_fwd(hidden.data_ptr(), w.data_ptr(), labels.data_ptr<int64_t>(), static_cast<float>(inv_tau), T,
                          V, H, hidden.stride(0), w.stride(0), scratch.data_ptr<float>(), S, lp.data_ptr<float>(),
                          ent.data_ptr<float>(), lse.data_ptr<float>(), ez.data_ptr<float>(), cur_stream()),
        "linear_ce_fwd");
  return {lp, ent, lse, ez};
}

at::Tensor linear_ce_dz(const at::Tensor& hidden, const at::Tensor& w_slice, const at::Tensor& labels,
                        const at::Tensor& lse, const at::Tensor& ez, const c10::optional<at::Tensor>& g_lp,
                        const c10::optional<at::Tensor>& g_ent, double inv_tau, int64_t v0) {
  CHECK_BF16(hidden); CHECK_BF16(w_slice); CHECK_F32(lse); CHECK_F32(ez);
  const int T = hidden.size(0), H = hidden.size(1), Vc = w_slice.size(0);
  TORCH_CHECK(Vc % 8 == 0);
  c10::cuda::CUDAGuard g(hidden.device());
  at::Tensor dz = at::empty({T, Vc}, hidden.options());
  const float* glp = g_lp.has_value() ? g_lp->data_ptr<float>() : nullptr;
  const float* gent = g_ent.has_value() ? g_ent->data_ptr<float>() : nullptr;
  check(prl_linear_ce_dz(hidden.data_ptr(), w_slice.data_ptr(), labels.data_ptr<int64_t>(), lse.data_ptr<float>(),
                         ez.data_ptr<float>(), glp, gent, static_cast<float>(inv_tau), T, Vc, H, static_cast<int>(v0),
                         hidden.stride(0), w_slice.stride(0), dz.data_ptr(), dz.stride(0), cur_stream()),
        "linear_ce_dz");
  return dz;
}

std::vector<at::Tensor> rmsnorm_fwd(const at::Tensor& x, const at::Tensor& w, double eps) {
  CHECK_BF16(x); CHECK_BF16(w);
  TORCH_CHECK(x.is_contiguous() && w.is_contiguous());
  const int H = x.size(-1);
  const int T = x.numel() / H;
  c10::cuda::CUDAGuard g(x.device());
  at::Tensor y = at::empty_like(x);
  at::Tensor rstd = at::empty({T}, x.options().dtype(at::kFloat));
  check(prl_rmsnorm_fwd(x.data_ptr(), w.data_ptr(), y.data_ptr(), rstd.data_ptr<float>(), T, H,
                        static_cast<float>(eps), cur_stream()), "rmsnorm_fwd");
  return {y, rstd};
}

// x += r (in place), returns rmsnorm(x) * w  — decode-path fusion of the residual add into the following norm.
at::Tensor add_rmsnorm_(at::Tensor x, const at::Tensor& r, const at::Tensor& w, double eps) {
  CHECK_BF16(x); CHECK_BF16(r); CHECK_BF16(w);
  TORCH_CHECK(x.is_contiguous() && r.is_contiguous() && x.sizes() == r.sizes());
  const int H = x.size(-1);
  const int T = x.numel() / H;
  c10::cuda::CUDAGuard g(x.device());
  at::Tensor y = at::empty_like(x);
  check(prl_add_rmsnorm(x.data_ptr(), r.data_ptr(), w.data_ptr(), y.data_ptr(), T, H, static_cast<float>(eps), cur_stream()), "add_rmsnorm");
  return y;
}

std::vector<at::Tensor> rmsnorm_bwd(const at::Tensor& x, const at::Tensor& w, const at::Tensor& dy,
                                    const at::Tensor& rstd) {
  CHECK_BF16(x); CHECK_BF16(w); CHECK_BF16(dy); CHECK_F32(rstd);
  TORCH_CHECK(x.is_contiguous() && dy.is_contiguous());
  const int H = x.size(-1);
  const int T = x.numel() / H;
  c10::cuda::CUDAGuard g(x.device());
  at::Tensor dx = at::empty_like(x);
  at::Tensor dw = at::zeros({H}, x.options().dtype(at::kFloat));
  check(prl_rmsnorm_bwd(x.data_ptr(), w.data_ptr(), dy.data_ptr(), rstd.data_ptr<float>(), dx.data_ptr(),
                        dw.data_ptr<float>(), T, H, cur_stream()), "rmsnorm_bwd");
  return {dx, dw};
}

void rope_(at::Tensor qkv, const at::Tensor& pos, const at::Tensor& cos_tab, const at::Tensor& sin_tab, int64_t nq,
           int64_t nkv, int64_t d, double sign, const c10::optional<at::Tensor>& k_cache,
           const c10::optional<at::Tensor>& v_cache, const c10::optional<at::Tensor>& slot_mapping,
           int64_t block_size) {
  CHECK_BF16(qkv); CHECK_F32(cos_tab); CHECK_F32(sin_tab);
  TORCH_CHECK(pos.is_cuda() && pos.scalar_type() == at::kLong && pos.is_contiguous());
  TORCH_CHECK(qkv.dim() == 2 && qkv.stride(1) == 1);
  c10::cuda::CUDAGuard g(qkv.device());
  void* kc = nullptr; void* vc = nullptr; const int64_t* sm = nullptr;
  if (slot_mapping.has_value()) {
    TORCH_CHECK(k_cache.has_value() && v_cache.has_value());
    kc = k_cache->data_ptr(); vc = v_cache->data_ptr();
    TORCH_CHECK(slot_mapping->scalar_type() == at::kLong);
    sm = slot_mapping->data_ptr<int64_t>();
  }
  check(prl_rope(qkv.data_ptr(), pos.data_ptr<int64_t>(), cos_tab.data_ptr<float>(), sin_tab.data_ptr<float>(),
                 qkv.size(0), nq, nkv, d, qkv.stride(0), static_cast<float>(sign), kc, vc, sm, block_size,
                 cur_stream()), "rope");
}

at::Tensor swiglu_bwd(const at::Tensor& gu, const at::Tensor& dh) {
  CHECK_BF16(gu); CHECK_BF16(dh);
  TORCH_CHECK(gu.is_contiguous() && dh.is_contiguous());
  const int F = dh.size(-1);
  const int64_t T = dh.numel() / F;
  c10::cuda::CUDAGuard g(gu.device());
  at::Tensor dgu = at::empty_like(gu);
  check(prl_swiglu_bwd(gu.data_ptr(), dh.data_ptr(), dgu.data_ptr(), T, F, cur_stream()), "swiglu_bwd");
  return dgu;
}

std::vector<at::Tensor> quant_fp8_rowwise(const at::Tensor& x) {
  CHECK_BF16(x); TORCH_CHECK(x.is_contiguous() && x.dim() == 2);
  c10::cuda::CUDAGuard g(x.device());
  at::Tensor q = at::empty(x.sizes(), x.options().dtype(at::kFloat8_e4m3fn));
  at::Tensor s = at::empty({x.size(0)}, x.options().dtype(at::kFloat));
  check(prl_quant_fp8_rowwise(x.data_ptr(), q.data_ptr(), s.data_ptr<float>(), x.size(0), x.size(1), cur_stream()),
        "quant_fp8_rowwise");
  return {q, s};
}

void grad_accum_(at::Tensor acc, const at::Tensor& g, double scale, const c10::optional<at::Tensor>& sumsq) {
  CHECK_F32(acc); TORCH_CHECK(acc.is_contiguous() && g.is_contiguous() && acc.numel() == g.numel());
  const int dt = g.scalar_type() == at::kFloat ? 0 : 1;
  TORCH_CHECK(dt == 0 || g.scalar_type() == at::kBFloat16);
  c10::cuda::CUDAGuard gd(acc.device());
  check(prl_grad_accum(acc.data_ptr<float>(), g.data_ptr(), dt, static_cast<float>(scale), acc.numel(),
                       sumsq.has_value() ? sumsq->data_ptr<float>() : nullptr, cur_stream()), "grad_accum");
}

void sumsq_(const at::Tensor& g, at::Tensor out) {
  CHECK_F32(g); CHECK_F32(out);
  c10::cuda::CUDAGuard gd(g.device());
  check(prl_sumsq(g.data_ptr<float>(), g.numel(), out.data_ptr<float>(), cur_stream()), "sumsq");
}

void adamw_(at::Tensor master, at::Tensor m, at::Tensor v, at::Tensor grad, const c10::optional<at::Tensor>& param_bf16,
            double lr, double beta1, double beta2, double eps, double wd, int64_t step, double max_norm,
            const c10::optional<at::Tensor>& total_sumsq, const c10::optional<at::Tensor>& status, bool zero_grad) {
  CHECK_F32(master); CHECK_F32(m); CHECK_F32(v); CHECK_F32(grad);
  TORCH_CHECK(master.is_contiguous() && m.is_contiguous() && v.is_contiguous() && grad.is_contiguous());
  c10::cuda::CUDAGuard gd(master.device());
  void* pb = nullptr;
  if (param_bf16.has_value()) { CHECK_BF16(*param_bf16); TORCH_CHECK(param_bf16->is_contiguous()); pb = param_bf16->data_ptr(); }
  check(prl_adamw(master.data_ptr<float>(), m.data_ptr<float>(), v.data_ptr<float>(), grad.data_ptr<float>(), pb,
                  master.numel(), lr, beta1, beta2, eps, wd, static_cast<int>(step), max_norm,
                  total_sumsq.has_value() ? total_sumsq->data_ptr<float>() : nullptr,
                  status.has_value() ? status->data_ptr<int>() : nullptr, zero_grad ? 1 : 0, cur_stream()), "adamw");
}

at::Tensor paged_decode(const at::Tensor& q, const at::Tensor& k_cache, const at::Tensor& v_cache,
                        const at::Tensor& block_table, const at::Tensor& seq_lens, int64_t nq, int64_t nkv, int64_t d,
                        double sm_scale, int64_t splits, c10::optional<at::Tensor> out) {
  CHECK_BF16(q); CHECK_BF16(k_cache); CHECK_BF16(v_cache);
  TORCH_CHECK(block_table.scalar_type() == at::kInt && seq_lens.scalar_type() == at::kInt);
  TORCH_CHECK(q.dim() == 2 && q.stride(1) == 1 && k_cache.is_contiguous() && v_cache.is_contiguous());
  TORCH_CHECK(k_cache.size(2) == 16, "KV page size must be 16 tokens");
  const int B = q.size(0);
  c10::cuda::CUDAGuard g(q.device());
  at::Tensor o = out.has_value() ? *out : at::empty({B, nq * d}, q.options());
  const int64_t ws = prl_paged_decode_workspace_floats(B, nq, d, splits);
  at::Tensor workspace = at::empty({ws}, q.options().dtype(at::kFloat));
  check(prl_paged_decode(q.data_ptr(), q.stride(0), k_cache.data_ptr(), v_cache.data_ptr(),
                         block_table.data_ptr<int32_t>(), block_table.stride(0), seq_lens.data_ptr<int32_t>(), B, nq,
                         nkv, d, static_cast<float>(sm_scale), splits, workspace.data_ptr<float>(), o.data_ptr(),
                         o.stride(0), cur_stream()), "paged_decode");
  return o;
}

std::vector<at::Tensor> sample(const at::Tensor& logits, const c10::optional<at::Tensor>& inv_temp,
                               const c10::optional<at::Tensor>& seeds, bool greedy) {
  TORCH_CHECK(logits.is_cuda() && logits.dim() == 2 && logits.stride(1) == 1);
  const int dt = logits.scalar_type() == at::kFloat ? 0 : 1;
  TORCH_CHECK(dt == 0 || logits.scalar_type() == at::kBFloat16);
  const int B = logits.size(0), V = logits.size(1);
  c10::cuda::CUDAGuard g(logits.device());
  at::Tensor tok = at::empty({B}, logits.options().dtype(at::kLong));
  at::Tensor lp = at::empty({B}, logits.options().dtype(at::kFloat));
  const uint64_t* sd = nullptr;
  if (seeds.has_value()) { TORCH_CHECK(seeds->scalar_type() == at::kLong); sd = reinterpret_cast<const uint64_t*>(seeds->data_ptr<int64_t>()); }
  check(prl_sample(logits.data_ptr(), dt, logits.stride(0), B, V,
                   inv_temp.has_value() ? inv_temp->data_ptr<float>() : nullptr, sd, greedy ? 1 : 0,
                   tok.data_ptr<int64_t>(), lp.data_ptr<float>(), cur_stream()), "sample");
  return {tok, lp};
}

std::vector<at::Tensor> sample_topk_topp(const at::Tensor& logits, const c10::optional<at::Tensor>& inv_temp,
                                         const c10::optional<at::Tensor>& seeds, bool greedy, int64_t top_k, double top_p) {
  TORCH_CHECK(logits.is_cuda() && logits.dim() == 2 && logits.stride(1) == 1);
  const int dt = logits.scalar_type() == at::kFloat ? 0 : 1;
  TORCH_CHECK(dt == 0 || logits.scalar_type() == at::kBFloat16);
  const int B = logits.size(0), V = logits.size(1);
  c10::cuda::CUDAGuard g(logits.device());
  at::Tensor tok = at::empty({B}, logits.options().dtype(at::kLong));
  at::Tensor lp = at::empty({B}, logits.options().dtype(at::kFloat));
  const uint64_t* sd = nullptr;
  if (seeds.has_value()) { TORCH_CHECK(seeds->scalar_type() == at::kLong); sd = reinterpret_cast<const uint64_t*>(seeds->data_ptr<int64_t>()); }
  check(prl_sample_topk_topp(logits.data_ptr(), dt, logits.stride(0), B, V,
                             inv_temp.has_value() ? inv_temp->data_ptr<float>() : nullptr, sd, greedy ? 1 : 0,
                             static_cast<int>(top_k), static_cast<float>(top_p), tok.data_ptr<int64_t>(), lp.data_ptr<float>(),
                             cur_stream()), "sample_topk_topp");
  return {tok, lp};
}

// varlen causal flash-attention forward on tcgen05 (csrc/attn_fwd.cu).  q:[Tq,nq,128] k,v:[Tk,nkv,128] (row-strided views allowed)
std::vector<at::Tensor> attn_fwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& cu_q,
                                 const at::Tensor& cu_k, double sm_scale, bool causal, bool want_lse,
                                 const c10::optional<at::Tensor>& dbg) {
  CHECK_BF16(q); CHECK_BF16(k); CHECK_BF16(v);
  TORCH_CHECK(q.dim() == 3 && k.dim() == 3 && v.dim() == 3 && q.size(2) == 128 && k.size(2) == 128 && v.size(2) == 128,
              "attn_fwd: head_dim must be 128");
  TORCH_CHECK(q.stride(2) == 1 && q.stride(1) == 128 && k.stride(2) == 1 && k.stride(1) == 128 && v.stride(2) == 1 && v.stride(1) == 128,
              "attn_fwd: heads must be contiguous inside a token row");
  TORCH_CHECK(q.stride(0) % 8 == 0 && k.stride(0) % 8 == 0 && v.stride(0) % 8 == 0);
  TORCH_CHECK((reinterpret_cast<uintptr_t>(q.data_ptr()) & 15) == 0 && (reinterpret_cast<uintptr_t>(k.data_ptr()) & 15) == 0 &&
              (reinterpret_cast<uintptr_t>(v.data_ptr()) & 15) == 0, "attn_fwd: 16-byte aligned operands");
  TORCH_CHECK(cu_q.is_cuda() && cu_q.scalar_type() == at::kInt && cu_q.is_contiguous() && cu_k.is_cuda() &&
              cu_k.scalar_type() == at::kInt && cu_k.is_contiguous() && cu_q.numel() == cu_k.numel());
  const int Tq = q.size(0), Tk = k.size(0), nq = q.size(1), nkv = k.size(1), B = cu_q.numel() - 1;
  c10::cuda::CUDAGuard g(q.device());
  at::Tensor o = at::empty({Tq, nq, 128}, q.options());
  at::Tensor lse = want_lse ? at::empty({nq, Tq}, q.options().dtype(at::kFloat)) : at::Tensor();
  check(prl_attn_fwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), want_lse ? lse.data_ptr<float>() : nullptr,
                     cu_q.data_ptr<int>(), cu_k.data_ptr<int>(), B, Tq, Tk, nq, nkv, q.stride(0), k.stride(0), v.stride(0),
                     static_cast<int64_t>(nq) * 128, static_cast<float>(sm_scale), causal ? 1 : 0, Tq / 128 + B,
                     dbg.has_value() ? reinterpret_cast<long long*>(dbg->data_ptr<int64_t>()) : nullptr, cur_stream()),
        "attn_fwd");
  if (want_lse) return {o, lse};
  return {o};
}

// flash-attention backward (csrc/attn_bwd.cu): returns (dq, dk, dv) in bf16, shapes of q / k / v (contiguous)
std::vector<at::Tensor> attn_bwd(const at::Tensor& q, const at::Tensor& k, const at::Tensor& v, const at::Tensor& o,
                                 const at::Tensor& d_o, const at::Tensor& lse, const at::Tensor& cu_q, const at::Tensor& cu_k,
                                 double sm_scale, bool causal) {
  CHECK_BF16(q); CHECK_BF16(k); CHECK_BF16(v); CHECK_BF16(o); CHECK_BF16(d_o); CHECK_F32(lse);
  for (const at::Tensor* t : {&q, &k, &v, &o, &d_o}) {
    TORCH_CHECK(t->dim() == 3 && t->size(2) == 128 && t->stride(2) == 1 && t->stride(1) == 128 && t->stride(0) % 8 == 0 &&
                (reinterpret_cast<uintptr_t>(t->data_ptr()) & 15) == 0, "attn_bwd: [T, heads, 128] row-strided bf16 operands");
  }
  const int Tq = q.size(0), Tk = k.size(0), nq = q.size(1), nkv = k.size(1), B = cu_q.numel() - 1;
  TORCH_CHECK(lse.is_contiguous() && lse.size(0) == nq && lse.size(1) == Tq);
  c10::cuda::CUDAGuard g(q.device());
  at::Tensor dq_acc = at::zeros({Tq, nq, 128}, q.options().dtype(at::kFloat));
  at::Tensor dk = at::empty({Tk, nkv, 128}, q.options());
  at::Tensor dv = at::empty({Tk, nkv, 128}, q.options());
  at::Tensor dsum = at::empty({nq, Tq}, q.options().dtype(at::kFloat));
  check(prl_attn_bwd(q.data_ptr(), k.data_ptr(), v.data_ptr(), o.data_ptr(), d_o.data_ptr(), lse.data_ptr<float>(),
                     dsum.data_ptr<float>(), dq_acc.data_ptr<float>(), dk.data_ptr(), dv.data_ptr(), cu_q.data_ptr<int>(),
                     cu_k.data_ptr<int>(), B, Tq, Tk, nq, nkv, q.stride(0), k.stride(0), v.stride(0), o.stride(0), d_o.stride(0),
                     static_cast<float>(sm_scale), causal ? 1 : 0, Tk / 128 + B, cur_stream()), "attn_bwd");
  return {dq_acc.to(at::kBFloat16), dk, dv};
}

// Decode-shape GEMM (swap-AB + split-K).  mode 0: out = x @ w.T (+bias)(+residual); mode 1: SwiGLU on interleaved W.
at::Tensor skinny_gemm(const at::Tensor& x, const at::Tensor& w, const c10::optional<at::Tensor>& residual,
                       const c10::optional<at::Tensor>& bias, const c10::optional<at::Tensor>& x_scale,
                       const c10::optional<at::Tensor>& w_scale, at::Tensor workspace, at::Tensor counters, int64_t mode,
                       int64_t splits) {
  TORCH_CHECK(x.is_cuda() && w.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.stride(1) == 1 && w.stride(1) == 1);
  const bool fp8 = x.element_size() == 1;
  TORCH_CHECK(fp8 == (w.element_size() == 1), "x and w must both be bf16 or both fp8");
  if (!fp8) { CHECK_BF16(x); CHECK_BF16(w); }
  const int B = x.size(0), K = x.size(1), N = w.size(0);
  TORCH_CHECK(w.size(1) == K && B <= 256 && K % 16 == 0);
  CHECK_F32(workspace);
  TORCH_CHECK(counters.scalar_type() == at::kInt && counters.numel() >= (N + 127) / 128);
  c10::cuda::CUDAGuard g(x.device());
  const int n_out = mode == 1 ? N / 2 : N;
  at::Tensor out = at::empty({B, n_out}, x.options().dtype(at::kBFloat16));
  const void* r = nullptr;
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->is_contiguous() && residual->numel() == out.numel()); r = residual->data_ptr(); }
  const void* bs = nullptr;
  if (bias.has_value()) { CHECK_BF16(*bias); bs = bias->data_ptr(); }
  const float* xs = x_scale.has_value() ? x_scale->data_ptr<float>() : nullptr;
  const float* ws = w_scale.has_value() ? w_scale->data_ptr<float>() : nullptr;
  TORCH_CHECK(!fp8 || (xs && ws), "fp8 skinny GEMM needs x_scale and w_scale");
  const int64_t need = prl_skinny_workspace_floats(B, N, splits > 0 ? splits : 16);
  TORCH_CHECK(splits == 1 || workspace.numel() >= need || workspace.numel() >= (int64_t)(148 + (N + 127) / 128) * 128 * (((B + 15) / 16) * 16),
              "skinny GEMM workspace too small");
  check(prl_skinny_gemm(x.data_ptr(), w.data_ptr(), out.data_ptr(), r, bs, xs, ws, workspace.data_ptr<float>(), counters.data_ptr<int>(),
                        B, N, K, x.stride(0), w.stride(0), out.stride(0), static_cast<int>(mode), fp8 ? 1 : 0,
                        static_cast<int>(splits), cur_stream()), "skinny_gemm");
  return out;
}

// Decode-shape GEMM v2: streamed weights, cluster split-K with a DSMEM reduce-scatter (csrc/decode_gemm.cu).
at::Tensor stream_gemm(const at::Tensor& x, const at::Tensor& w, const c10::optional<at::Tensor>& residual,
                       const c10::optional<at::Tensor>& bias, const c10::optional<at::Tensor>& x_scale,
                       const c10::optional<at::Tensor>& w_scale, int64_t mode, int64_t splits,
                       c10::optional<at::Tensor> out_opt, int64_t pair) {
  TORCH_CHECK(x.is_cuda() && w.is_cuda() && x.dim() == 2 && w.dim() == 2 && x.stride(1) == 1 && w.stride(1) == 1);
  const bool fp8 = x.element_size() == 1;
  TORCH_CHECK(fp8 == (w.element_size() == 1), "x and w must both be bf16 or both fp8");
  if (!fp8) { CHECK_BF16(x); CHECK_BF16(w); }
  const int B = x.size(0), K = x.size(1), N = w.size(0);
  TORCH_CHECK(w.size(1) == K && B <= 256 && K % 16 == 0);
  c10::cuda::CUDAGuard g(x.device());
  const int n_out = mode == 1 ? N / 2 : N;
  at::Tensor out = out_opt.has_value() ? *out_opt : at::empty({B, n_out}, x.options().dtype(at::kBFloat16));
  CHECK_BF16(out);
  TORCH_CHECK(out.dim() == 2 && out.size(0) == B && out.size(1) == n_out && out.stride(1) == 1);
  const void* r = nullptr;
  if (residual.has_value()) {
    CHECK_BF16(*residual);
    TORCH_CHECK(residual->dim() == 2 && residual->stride(1) == 1 && residual->stride(0) == out.stride(0) && residual->numel() == out.numel());
    r = residual->data_ptr();
  }
  const void* bs = nullptr;
  if (bias.has_value()) { CHECK_BF16(*bias); bs = bias->data_ptr(); }
  const float* xs = x_scale.has_value() ? x_scale->data_ptr<float>() : nullptr;
  const float* ws = w_scale.has_value() ? w_scale->data_ptr<float>() : nullptr;
  TORCH_CHECK(!fp8 || (xs && ws), "fp8 stream GEMM needs x_scale and w_scale");
  check(prl_stream_gemm(x.data_ptr(), w.data_ptr(), out.data_ptr(), r, bs, xs, ws, B, N, K, x.stride(0), w.stride(0), out.stride(0),
                        static_cast<int>(mode), fp8 ? 1 : 0, static_cast<int>(splits), static_cast<int>(pair), cur_stream()), "stream_gemm");
  return out;
}

// Row-parallel TP linear with the all-reduce fused into the GEMM epilogue over NVLink peer memory.
at::Tensor skinny_gemm_tp(const at::Tensor& x, const at::Tensor& w, const c10::optional<at::Tensor>& residual,
                          at::Tensor workspace, at::Tensor counters, int64_t splits, const std::vector<int64_t>& inbox_ptrs,
                          const std::vector<int64_t>& flag_ptrs, int64_t tp_rank, at::Tensor epoch_state, int64_t inbox_half) {
  CHECK_BF16(x); CHECK_BF16(w);
  TORCH_CHECK(x.dim() == 2 && w.dim() == 2 && x.stride(1) == 1 && w.stride(1) == 1 && x.size(1) == w.size(1));
  const int B = x.size(0), K = x.size(1), N = w.size(0);
  const int world = inbox_ptrs.size();
  const int BP = ((B + 15) / 16) * 16;
  TORCH_CHECK((int64_t)world * ((N + 127) / 128) * BP * 128 <= inbox_half, "TP inbox too small");
  TORCH_CHECK(epoch_state.is_cuda() && epoch_state.scalar_type() == at::kInt && epoch_state.numel() >= 2);
  c10::cuda::CUDAGuard g(x.device());
  at::Tensor out = at::empty({B, N}, x.options());
  const void* r = nullptr;
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->is_contiguous()); r = residual->data_ptr(); }
  auto ib = to_ptrs(inbox_ptrs);
  auto fl = to_ptrs(flag_ptrs);
  check(prl_skinny_gemm_tp(x.data_ptr(), w.data_ptr(), out.data_ptr(), r, nullptr, nullptr, nullptr, workspace.data_ptr<float>(),
                           counters.data_ptr<int>(), B, N, K, x.stride(0), w.stride(0), out.stride(0), 0, 0, static_cast<int>(splits),
                           world, static_cast<int>(tp_rank), reinterpret_cast<uint32_t*>(epoch_state.data_ptr<int>()), epoch_state.data_ptr<int>() + 1,
                           inbox_half, ib.data(), fl.data(), cur_stream()),
        "skinny_gemm_tp");
  return out;
}

// ---------------------------------------------------------------- symmetric heap / peer kernels
int64_t sym_malloc(int64_t bytes) {
  void* p = nullptr;
  check(prl_cuda_malloc(&p, bytes), "cuda_malloc");
  return reinterpret_cast<int64_t>(p);
}
void sym_free(int64_t ptr) { check(prl_cuda_free(reinterpret_cast<void*>(ptr)), "cuda_free"); }
py::bytes ipc_get_handle(int64_t ptr) {
  char buf[64];
  check(prl_ipc_get_handle(reinterpret_cast<void*>(ptr), buf), "ipc_get_handle");
  return py::bytes(buf, 64);
}
int64_t ipc_open_handle(const std::string& h) {
  TORCH_CHECK(h.size() == 64);
  void* p = nullptr;
  check(prl_ipc_open_handle(h.data(), &p), "ipc_open_handle");
  return reinterpret_cast<int64_t>(p);
}
void ipc_close_handle(int64_t ptr) { check(prl_ipc_close_handle(reinterpret_cast<void*>(ptr)), "ipc_close_handle"); }
void enable_peer_access(int64_t dev) { check(prl_enable_peer_access(static_cast<int>(dev)), "enable_peer_access"); }

// ---- CUDA VMM ranges (csrc/vmm.cu): fixed virtual address, physical backing released / re-created on demand
std::vector<int64_t> vmm_reserve(int64_t bytes, int64_t device) {
  int64_t ptr = 0, size = 0;
  check(prl_vmm_reserve(bytes, static_cast<int>(device), &ptr, &size), "vmm_reserve");
  return {ptr, size};
}
void vmm_map(int64_t ptr) { check(prl_vmm_map(ptr), "vmm_map (out of device memory?)"); }
void vmm_unmap(int64_t ptr) { check(prl_vmm_unmap(ptr), "vmm_unmap"); }
void vmm_free(int64_t ptr) { check(prl_vmm_free(ptr), "vmm_free"); }
bool vmm_is_mapped(int64_t ptr) { return prl_vmm_is_mapped(ptr) != 0; }

at::Tensor tensor_from_ptr(int64_t ptr, std::vector<int64_t> sizes, at::ScalarType dtype, int64_t device) {
  auto opts = at::TensorOptions().dtype(dtype).device(at::kCUDA, device);
  return at::from_blob(reinterpret_cast<void*>(ptr), sizes, [](void*) {}, opts);
}

std::vector<void*> to_ptrs(const std::vector<int64_t>& v) {
  std::vector<void*> r;
  for (auto x : v) r.push_back(reinterpret_cast<void*>(x));
  return r;
}

// epoch_ctr: int32 device tensor (1 element) owned by this rank; the kernel increments it (graph-replay safe).
void p2p_barrier(const std::vector<int64_t>& flag_ptrs, int64_t rank, at::Tensor epoch_ctr) {
  TORCH_CHECK(epoch_ctr.is_cuda() && epoch_ctr.scalar_type() == at::kInt && epoch_ctr.numel() >= 1);
  c10::cuda::CUDAGuard g(epoch_ctr.device());
  auto p = to_ptrs(flag_ptrs);
  check(prl_p2p_barrier(p.data(), rank, p.size(), reinterpret_cast<uint32_t*>(epoch_ctr.data_ptr<int>()), cur_stream()), "p2p_barrier");
}
void p2p_copy(int64_t src, int64_t dst, int64_t nbytes, int64_t max_ctas) {
  check(prl_p2p_copy(reinterpret_cast<void*>(src), reinterpret_cast<void*>(dst), nbytes, max_ctas, cur_stream()), "p2p_copy");
}
void p2p_push_fp8(const at::Tensor& src, int64_t dst_q, int64_t dst_scale, int64_t max_ctas) {
  CHECK_BF16(src); TORCH_CHECK(src.is_contiguous() && src.dim() == 2);
  check(prl_p2p_push_fp8(src.data_ptr(), reinterpret_cast<void*>(dst_q), reinterpret_cast<float*>(dst_scale),
                         src.size(0), src.size(1), max_ctas, cur_stream()), "p2p_push_fp8");
}
// out = sum over ranks of bf16 partials living at buf_ptrs[r] + elem_off (+ residual); rank-ordered fp32 sum.
void p2p_allreduce(const std::vector<int64_t>& buf_ptrs, int64_t elem_off, at::Tensor out, const c10::optional<at::Tensor>& residual,
                   int64_t max_ctas) {
  CHECK_BF16(out);
  TORCH_CHECK(out.is_contiguous() && out.numel() % 8 == 0);
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->is_contiguous() && residual->numel() == out.numel()); }
  c10::cuda::CUDAGuard g(out.device());
  auto p = to_ptrs(buf_ptrs);
  check(prl_p2p_allreduce(p.data(), elem_off, out.data_ptr(), residual.has_value() ? residual->data_ptr() : nullptr, out.numel(),
                          static_cast<int>(p.size()), static_cast<int>(max_ctas), cur_stream()), "p2p_allreduce");
}

void p2p_allgather_pull(const std::vector<int64_t>& shard_ptrs, int64_t out, int64_t shard_bytes, int64_t rank,
                        int64_t max_ctas) {
  auto p = to_ptrs(shard_ptrs);
  check(prl_p2p_allgather_pull(p.data(), reinterpret_cast<void*>(out), shard_bytes, rank, p.size(), max_ctas,
                               cur_stream()), "p2p_allgather_pull");
}
void p2p_reduce_scatter_accum(const std::vector<int64_t>& grad_ptrs, int64_t shard_off, at::Tensor acc, double scale,
                              int64_t rank, const c10::optional<at::Tensor>& sumsq, int64_t max_ctas) {
  CHECK_F32(acc);
  auto p = to_ptrs(grad_ptrs);
  check(prl_p2p_reduce_scatter_accum(p.data(), shard_off, acc.data_ptr<float>(), acc.numel(), scale, p.size(), rank,
                                     sumsq.has_value() ? sumsq->data_ptr<float>() : nullptr, max_ctas, cur_stream()),
        "p2p_reduce_scatter_accum");
}

// All-gather fused GEMM: D = A · gather(W_shards)^T with the gather done by copy warps of the same kernel.
at::Tensor ag_gemm_bf16(const at::Tensor& a, const std::vector<int64_t>& shard_ptrs, int64_t gathered_ptr,
                        int64_t flags_ptr, int64_t epoch, int64_t rows_per_rank, int64_t rank, int64_t comm_chunks) {
  CHECK_BF16(a); CHECK_LASTDIM(a);
  const int M = a.size(0), K = a.size(1);
  const int world = shard_ptrs.size();
  const int N = rows_per_rank * world;
  c10::cuda::CUDAGuard g(a.device());
  at::Tensor d = at::empty({M, N}, a.options());
  auto p = to_ptrs(shard_ptrs);
  check(prl_ag_gemm_bf16(a.data_ptr(), p.data(), reinterpret_cast<void*>(gathered_ptr), d.data_ptr(),
                         reinterpret_cast<uint32_t*>(flags_ptr), static_cast<uint32_t>(epoch), M, N, K, a.stride(0),
                         d.stride(0), world, rank, comm_chunks, cur_stream()), "ag_gemm_bf16");
  return d;
}

}  // namespace

// ---- NVLS: multicast-mapped symmetric heap + multimem collectives (csrc/nvls.cu)
void nvls_allreduce(int64_t mc_in, int64_t mc_out, const c10::optional<at::Tensor>& residual, int64_t bytes,
                    const std::vector<int64_t>& flag_ptrs, int64_t rank, at::Tensor state, int64_t ctas) {
  TORCH_CHECK(state.is_cuda() && state.scalar_type() == at::kInt && state.numel() >= 3);
  c10::cuda::CUDAGuard g(state.device());
  std::vector<void*> fp = to_ptrs(flag_ptrs);
  const void* r = nullptr;
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->is_contiguous() && residual->numel() * 2 >= bytes); r = residual->data_ptr(); }
  check(prl_nvls_allreduce(reinterpret_cast<const void*>(mc_in), reinterpret_cast<void*>(mc_out), r, bytes, fp.data(),
                           static_cast<int>(rank), static_cast<int>(fp.size()), reinterpret_cast<uint32_t*>(state.data_ptr<int>()),
                           static_cast<int>(ctas), cur_stream()), "nvls_allreduce");
}

// one barrier + in-switch reduction of the whole buffer into a LOCAL output tensor (+ residual)
at::Tensor nvls_allreduce_oneshot(int64_t mc_in, at::Tensor out, const c10::optional<at::Tensor>& residual,
                                  const std::vector<int64_t>& flag_ptrs, int64_t rank, at::Tensor state, int64_t ctas) {
  CHECK_BF16(out);
  TORCH_CHECK(out.is_contiguous() && state.is_cuda() && state.scalar_type() == at::kInt && state.numel() >= 3);
  const int64_t bytes = out.numel() * 2;
  TORCH_CHECK(bytes % 16 == 0, "all-reduce size must be a multiple of 16 bytes");
  c10::cuda::CUDAGuard g(out.device());
  std::vector<void*> fp = to_ptrs(flag_ptrs);
  const void* r = nullptr;
  if (residual.has_value()) { CHECK_BF16(*residual); TORCH_CHECK(residual->is_contiguous() && residual->numel() == out.numel()); r = residual->data_ptr(); }
  check(prl_nvls_allreduce_oneshot(reinterpret_cast<const void*>(mc_in), out.data_ptr(), r, bytes, fp.data(), static_cast<int>(rank),
                                   static_cast<int>(fp.size()), reinterpret_cast<uint32_t*>(state.data_ptr<int>()),
                                   static_cast<int>(ctas), cur_stream()), "nvls_allreduce_oneshot");
  return out;
}

void nvls_reduce_scatter_accum(int64_t mc_grads, int64_t shard_off_elems, at::Tensor acc, double scale,
                               const c10::optional<at::Tensor>& sumsq, int64_t ctas) {
  CHECK_F32(acc);
  TORCH_CHECK(acc.is_contiguous());
  c10::cuda::CUDAGuard g(acc.device());
  float* ss = sumsq.has_value() ? sumsq->data_ptr<float>() : nullptr;
  check(prl_nvls_reduce_scatter_accum(reinterpret_cast<const void*>(mc_grads), shard_off_elems, acc.data_ptr<float>(), acc.numel(),
                                      static_cast<float>(scale), ss, static_cast<int>(ctas), cur_stream()), "nvls_reduce_scatter_accum");
}

void nvls_bcast_store(const at::Tensor& src, int64_t mc_dst, int64_t ctas) {
  TORCH_CHECK(src.is_cuda() && src.is_contiguous());
  c10::cuda::CUDAGuard g(src.device());
  check(prl_nvls_bcast_store(src.data_ptr(), reinterpret_cast<void*>(mc_dst), src.numel() * src.element_size(), static_cast<int>(ctas),
                             cur_stream()), "nvls_bcast_store");
}

static at::Tensor vec_i64(const std::vector<int64_t>& v) {
  return at::tensor(c10::ArrayRef<int64_t>(v.data(), v.size()), at::TensorOptions().dtype(at::kLong));
}

PYBIND11_MODULE(TORCH_EXTENSION_NAME, m) {
  py::class_<prl::EngineScheduler>(m, "EngineScheduler")
      .def(py::init<prl::KvPageManager*, int64_t, int64_t, bool>(), py::arg("kv"), py::arg("max_num_seqs"), py::arg("page_size"),
           py::arg("prefix_cache"), py::keep_alive<1, 2>())
      .def("add", &prl::EngineScheduler::add, py::arg("prompt"), py::arg("max_new"), py::arg("group"), py::arg("cacheable"))
      .def("remove_waiting", &prl::EngineScheduler::remove_waiting)
      .def("num_waiting", &prl::EngineScheduler::num_waiting)
      .def("num_free_slots", &prl::EngineScheduler::num_free_slots)
      .def("num_live", &prl::EngineScheduler::num_live)
      .def("admit", &prl::EngineScheduler::admit)
      .def("info", [](const prl::EngineScheduler& s, int64_t h) {        // (slot, blocks, shared_prefix_blocks, cached_tokens, donor)
        const prl::SchedRequest& r = s.get(h);
        return py::make_tuple(r.slot, r.blocks, r.shared_prefix_blocks, r.cached_tokens, r.donor);
      })
      .def("prefill_arrays", [](const prl::EngineScheduler& s, const std::vector<std::tuple<int64_t, int64_t, int64_t>>& spans, bool gather) {
        const prl::PrefillArrays a = s.prefill_arrays(spans, gather);
        return py::make_tuple(vec_i64(a.ids), vec_i64(a.pos), vec_i64(a.slots), vec_i64(a.lens), vec_i64(a.g_blk), vec_i64(a.g_off), vec_i64(a.klens));
      })
      .def("block_rows", [](const prl::EngineScheduler& s, const std::vector<int64_t>& handles, int64_t width) {
        const std::vector<int32_t> rows = s.block_rows(handles, width);
        return at::tensor(c10::ArrayRef<int32_t>(rows.data(), rows.size()), at::TensorOptions().dtype(at::kInt))
            .reshape({static_cast<int64_t>(handles.size()), width});
      })
      .def("move_slots", &prl::EngineScheduler::move_slots)
      .def("free_slots", &prl::EngineScheduler::free_slots)
      .def("finish", &prl::EngineScheduler::finish, py::arg("handle"), py::arg("tokens"), py::arg("to_cache"));
  py::class_<prl::KvPageManager>(m, "KvPageManager")
      .def(py::init<int64_t, int64_t>(), py::arg("num_blocks"), py::arg("page_size"))
      .def_property_readonly("num_blocks", &prl::KvPageManager::num_blocks)
      .def("available", &prl::KvPageManager::available)
      .def("ref", &prl::KvPageManager::ref)
      .def("alloc", &prl::KvPageManager::alloc)
      .def("share", &prl::KvPageManager::share)
      .def("release", &prl::KvPageManager::release)
      .def("page_keys", &prl::KvPageManager::page_keys)
      .def("match", &prl::KvPageManager::match)
      .def("insert", &prl::KvPageManager::insert)
      .def("evict", &prl::KvPageManager::evict)
      .def("flush", &prl::KvPageManager::flush)
      .def("cached_pages", &prl::KvPageManager::cached_pages)
      .def_property_readonly("hits", &prl::KvPageManager::hits)
      .def_property_readonly("misses", &prl::KvPageManager::misses);
  m.def("gemm_bf16", &gemm_bf16, py::arg("a"), py::arg("b"), py::arg("residual") = py::none(),
        py::arg("out") = py::none(), py::arg("force_bn") = 0);
  m.def("gemm2_bf16", &gemm2_bf16, py::arg("a"), py::arg("b"), py::arg("a_mn") = false, py::arg("b_mn") = false,
        py::arg("residual") = py::none(), py::arg("out") = py::none(), py::arg("force_bn") = 0);
  m.def("gemm2_accum_f32_", &gemm2_accum_f32_);
  m.def("gemm_fp8", &gemm_fp8, py::arg("a"), py::arg("b"), py::arg("row_scale"), py::arg("col_scale"),
        py::arg("force_bn") = 0);
  m.def("gemm_swiglu", &gemm_swiglu);
  m.def("skinny_gemm", &skinny_gemm);
  m.def("skinny_gemm_tp", &skinny_gemm_tp);
  m.def("stream_gemm", &stream_gemm, py::arg("x"), py::arg("w"), py::arg("residual") = py::none(), py::arg("bias") = py::none(),
        py::arg("x_scale") = py::none(), py::arg("w_scale") = py::none(), py::arg("mode") = 0, py::arg("splits") = 0,
        py::arg("out") = py::none(), py::arg("pair") = -1);
  m.def("linear_ce_fwd", &linear_ce_fwd);
  m.def("linear_ce_dz", &linear_ce_dz);
  m.def("rmsnorm_fwd", &rmsnorm_fwd);
  m.def("rmsnorm_bwd", &rmsnorm_bwd);
  m.def("add_rmsnorm_", &add_rmsnorm_);
  m.def("rope_", &rope_);
  m.def("swiglu_bwd", &swiglu_bwd);
  m.def("quant_fp8_rowwise", &quant_fp8_rowwise);
  m.def("grad_accum_", &grad_accum_);
  m.def("sumsq_", &sumsq_);
  m.def("adamw_", &adamw_);
  m.def("paged_decode", &paged_decode, py::arg("q"), py::arg("k_cache"), py::arg("v_cache"), py::arg("block_table"),
        py::arg("seq_lens"), py::arg("nq"), py::arg("nkv"), py::arg("d"), py::arg("sm_scale"), py::arg("splits") = 1,
        py::arg("out") = py::none());
  m.def("attn_fwd", &attn_fwd, py::arg("q"), py::arg("k"), py::arg("v"), py::arg("cu_q"), py::arg("cu_k"), py::arg("sm_scale"),
        py::arg("causal") = true, py::arg("want_lse") = true, py::arg("dbg") = py::none());
  m.def("attn_bwd", &attn_bwd);
  m.def("sample", &sample);
  m.def("sample_topk_topp", &sample_topk_topp);
  m.def("sym_malloc", &sym_malloc);
  m.def("sym_free", &sym_free);
  m.def("ipc_get_handle", &ipc_get_handle);
  m.def("ipc_open_handle", &ipc_open_handle);
  m.def("ipc_close_handle", &ipc_close_handle);
  m.def("enable_peer_access", &enable_peer_access);
  m.def("tensor_from_ptr", &tensor_from_ptr);
  m.def("vmm_reserve", &vmm_reserve);
  m.def("vmm_map", &vmm_map);
  m.def("vmm_unmap", &vmm_unmap);
  m.def("vmm_free", &vmm_free);
  m.def("vmm_is_mapped", &vmm_is_mapped);
  m.def("vmm_granularity", [](int64_t d) { return prl_vmm_granularity(static_cast<int>(d)); });
  m.def("nvls_supported", [](int64_t d) { return prl_nvls_supported(static_cast<int>(d)); });
  m.def("nvls_round_size", [](int64_t b, int64_t d, int64_t w) { return prl_nvls_round_size(b, static_cast<int>(d), static_cast<int>(w)); });
  m.def("nvls_create", [](int64_t b, int64_t d) { return prl_nvls_create(b, static_cast<int>(d)); });
  m.def("nvls_mc_create", [](int64_t b, int64_t w) { return prl_nvls_mc_create(b, static_cast<int>(w)); });
  m.def("nvls_export_fd", [](int64_t id) { return prl_nvls_export_fd(id); });
  m.def("nvls_import_fd", [](int64_t fd, int64_t b, int64_t d) { return prl_nvls_import_fd(static_cast<int>(fd), b, static_cast<int>(d)); });
  m.def("nvls_mc_add_device", [](int64_t id, int64_t d) { return prl_nvls_mc_add_device(id, static_cast<int>(d)); });
  m.def("nvls_mc_bind", [](int64_t mc, int64_t mem) { return prl_nvls_mc_bind(mc, mem); });
  m.def("nvls_map", [](int64_t id, int64_t d) { return prl_nvls_map(id, static_cast<int>(d)); });
  m.def("nvls_unmap", [](int64_t va) { return prl_nvls_unmap(va); });
  m.def("nvls_release", [](int64_t id) { return prl_nvls_release(id); });
  m.def("close_fd", [](int64_t fd) { return prl_close_fd(static_cast<int>(fd)); });
  m.def("nvls_allreduce", &nvls_allreduce);
  m.def("nvls_allreduce_oneshot", &nvls_allreduce_oneshot);
  m.def("nvls_reduce_scatter_accum", &nvls_reduce_scatter_accum);
  m.def("nvls_bcast_store", &nvls_bcast_store);
  m.def("p2p_barrier", &p2p_barrier);
  m.def("p2p_copy", &p2p_copy);
  m.def("p2p_push_fp8", &p2p_push_fp8);
  m.def("p2p_allgather_pull", &p2p_allgather_pull);
  m.def("p2p_allreduce", &p2p_allreduce);
  m.def("p2p_reduce_scatter_accum", &p2p_reduce_scatter_accum);
  m.def("ag_gemm_bf16", &ag_gemm_bf16);
}
