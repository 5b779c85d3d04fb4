// polyrl_b200 — NVLS: symmetric heap on the CUDA VMM API with an NVSwitch MULTICAST mapping, and the
// `multimem.*` collectives that reduce inside the switch.
//
// Round 1's symmetric heap was cudaMalloc + legacy CUDA IPC, which cannot back a multicast object
// (SURVEY §5.8).  Here every rank creates its share with cuMemCreate (exportable as a POSIX
// file descriptor), maps every peer's share for plain ld/st (same capability as before) AND binds its
// share to one multicast object whose mapping `mc` addresses "the same offset on every GPU":
//   multimem.ld_reduce [mc + off]   -> the switch reads all N copies and returns their sum (fp32 accumulate)
//   multimem.st        [mc + off]   -> the switch writes the value into all N copies
// so an all-reduce moves 1/N of the data per GPU through one load + one store instead of N-1 peer reads.
//
// The reference gets NVLS only through NCCL inside SGLang's TP all-reduce
// (3rdparty/verl/verl/workers/rollout/sglang_rollout/sglang_rollout.py:99 sets NCCL_NVLS_ENABLE); here the
// collective is ours: one-shot all-reduce (+ fused residual) for TP decode activations, reduce-scatter fused
// with the fp32 accumulate for ZeRO gradients, and the multicast all-gather / broadcast store.
//
// Driver entry points are resolved through the runtime (no link-time dependency on libcuda).
#include <cuda.h>
#include <cuda_runtime.h>
#include <unistd.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>

#include "common.cuh"

namespace {

template <typename Fn>
Fn drv(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
  return reinterpret_cast<Fn>(p);
}

using FnGran = CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
using FnReserve = CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
using FnAddrFree = CUresult (*)(CUdeviceptr, size_t);
using FnCreate = CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
using FnRelease = CUresult (*)(CUmemGenericAllocationHandle);
using FnMap = CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
using FnUnmap = CUresult (*)(CUdeviceptr, size_t);
using FnAccess = CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);
using FnExport = CUresult (*)(void*, CUmemGenericAllocationHandle, CUmemAllocationHandleType, unsigned long long);
using FnImport = CUresult (*)(CUmemGenericAllocationHandle*, void*, CUmemAllocationHandleType);
using FnDevAttr = CUresult (*)(int*, CUdevice_attribute, CUdevice);
using FnDevGet = CUresult (*)(CUdevice*, int);
using FnMcCreate = CUresult (*)(CUmemGenericAllocationHandle*, const CUmulticastObjectProp*);
using FnMcAddDev = CUresult (*)(CUmemGenericAllocationHandle, CUdevice);
using FnMcBindMem = CUresult (*)(CUmemGenericAllocationHandle, size_t, CUmemGenericAllocationHandle, size_t, size_t, unsigned long long);
using FnMcGran = CUresult (*)(size_t*, const CUmulticastObjectProp*, CUmulticastGranularity_flags);
using FnMcUnbind = CUresult (*)(CUmemGenericAllocationHandle, CUdevice, size_t, size_t);

struct Handle {
  CUmemGenericAllocationHandle h = 0;
  size_t bytes = 0;
  int device = 0;
};
std::map<int64_t, Handle> g_handles;       // id -> physical / multicast handle
std::map<int64_t, size_t> g_maps;          // va -> bytes
int64_t g_next_id = 1;
std::mutex g_mu;

CUmemAllocationProp props(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  p.requestedHandleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  return p;
}

}  // namespace

// 1 if the device (and driver) support NVSwitch multicast objects.
extern "C" int prl_nvls_supported(int device) {
  auto attr = drv<FnDevAttr>("cuDeviceGetAttribute");
  auto dget = drv<FnDevGet>("cuDeviceGet");
  if (!attr || !dget || !drv<FnMcCreate>("cuMulticastCreate")) return 0;
  cudaSetDevice(device);
  cudaFree(nullptr);
  CUdevice d;
  if (dget(&d, device) != CUDA_SUCCESS) return 0;
  int v = 0;
  if (attr(&v, CU_DEVICE_ATTRIBUTE_MULTICAST_SUPPORTED, d) != CUDA_SUCCESS) return 0;
  return v;
}

// Allocation size that satisfies both the VMM and the multicast granularity for `world` devices.
extern "C" int64_t prl_nvls_round_size(int64_t bytes, int device, int world) {
  auto gran = drv<FnGran>("cuMemGetAllocationGranularity");
  if (!gran) return -1;
  size_t g = 0;
  CUmemAllocationProp p = props(device);
  if (gran(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) return -2;
  auto mcg = drv<FnMcGran>("cuMulticastGetGranularity");
  if (mcg && world > 1) {
    CUmulticastObjectProp mp{};
    mp.numDevices = world;
    mp.size = g;
    mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
    size_t mg = 0;
    if (mcg(&mg, &mp, CU_MULTICAST_GRANULARITY_RECOMMENDED) == CUDA_SUCCESS && mg > g) g = mg;
  }
  return static_cast<int64_t>((static_cast<size_t>(bytes) + g - 1) / g * g);
}

// Physical allocation on `device`, exportable as a POSIX fd.  Returns a handle id (> 0) or a negative error.
extern "C" int64_t prl_nvls_create(int64_t bytes, int device) {
  cudaSetDevice(device);
  cudaFree(nullptr);
  auto create = drv<FnCreate>("cuMemCreate");
  if (!create) return -80;
  CUmemAllocationProp p = props(device);
  Handle h;
  h.bytes = static_cast<size_t>(bytes);
  h.device = device;
  CUresult rc = create(&h.h, h.bytes, &p, 0);
  if (rc != CUDA_SUCCESS) return -81000 - static_cast<int>(rc);
  std::lock_guard<std::mutex> lk(g_mu);
  const int64_t id = g_next_id++;
  g_handles[id] = h;
  return id;
}

// Multicast object spanning `world` devices (created by ONE rank, imported by the others).
extern "C" int64_t prl_nvls_mc_create(int64_t bytes, int world) {
  auto mc = drv<FnMcCreate>("cuMulticastCreate");
  if (!mc) return -82;
  CUmulticastObjectProp mp{};
  mp.numDevices = world;
  mp.size = static_cast<size_t>(bytes);
  mp.handleTypes = CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR;
  Handle h;
  h.bytes = mp.size;
  CUresult rc = mc(&h.h, &mp);
  if (rc != CUDA_SUCCESS) return -83000 - static_cast<int>(rc);
  std::lock_guard<std::mutex> lk(g_mu);
  const int64_t id = g_next_id++;
  g_handles[id] = h;
  return id;
}

extern "C" int prl_nvls_export_fd(int64_t id) {
  auto exp = drv<FnExport>("cuMemExportToShareableHandle");
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(id);
  if (!exp || it == g_handles.end()) return -84;
  int fd = -1;
  CUresult rc = exp(&fd, it->second.h, CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR, 0);
  if (rc != CUDA_SUCCESS) return -85000 - static_cast<int>(rc);
  return fd;
}

extern "C" int64_t prl_nvls_import_fd(int fd, int64_t bytes, int device) {
  auto imp = drv<FnImport>("cuMemImportFromShareableHandle");
  if (!imp) return -86;
  cudaSetDevice(device);
  cudaFree(nullptr);
  Handle h;
  h.bytes = static_cast<size_t>(bytes);
  h.device = device;
  CUresult rc = imp(&h.h, reinterpret_cast<void*>(static_cast<uintptr_t>(fd)), CU_MEM_HANDLE_TYPE_POSIX_FILE_DESCRIPTOR);
  if (rc != CUDA_SUCCESS) return -87000 - static_cast<int>(rc);
  std::lock_guard<std::mutex> lk(g_mu);
  const int64_t id = g_next_id++;
  g_handles[id] = h;
  return id;
}

extern "C" int prl_nvls_mc_add_device(int64_t mc_id, int device) {
  auto add = drv<FnMcAddDev>("cuMulticastAddDevice");
  auto dget = drv<FnDevGet>("cuDeviceGet");
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(mc_id);
  if (!add || !dget || it == g_handles.end()) return -88;
  CUdevice d;
  if (dget(&d, device) != CUDA_SUCCESS) return -89;
  CUresult rc = add(it->second.h, d);
  return rc == CUDA_SUCCESS ? 0 : -90000 - static_cast<int>(rc);
}

// Bind this rank's physical allocation at offset 0 of the multicast object (after every device was added).
extern "C" int prl_nvls_mc_bind(int64_t mc_id, int64_t mem_id) {
  auto bind = drv<FnMcBindMem>("cuMulticastBindMem");
  std::lock_guard<std::mutex> lk(g_mu);
  auto mc = g_handles.find(mc_id), mem = g_handles.find(mem_id);
  if (!bind || mc == g_handles.end() || mem == g_handles.end()) return -91;
  CUresult rc = bind(mc->second.h, 0, mem->second.h, 0, mem->second.bytes, 0);
  return rc == CUDA_SUCCESS ? 0 : -92000 - static_cast<int>(rc);
}

// Map a (physical or multicast) handle into this process' address space with read/write access for `device`.
extern "C" int64_t prl_nvls_map(int64_t id, int device) {
  auto reserve = drv<FnReserve>("cuMemAddressReserve");
  auto map = drv<FnMap>("cuMemMap");
  auto access = drv<FnAccess>("cuMemSetAccess");
  auto afree = drv<FnAddrFree>("cuMemAddressFree");
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(id);
  if (!reserve || !map || !access || it == g_handles.end()) return -93;
  cudaSetDevice(device);
  const size_t bytes = it->second.bytes;
  CUdeviceptr va = 0;
  if (reserve(&va, bytes, 0, 0, 0) != CUDA_SUCCESS) return -94;
  CUresult rc = map(va, bytes, 0, it->second.h, 0);
  if (rc != CUDA_SUCCESS) {
    if (afree) afree(va, bytes);
    return -95000 - static_cast<int>(rc);
  }
  CUmemAccessDesc d{};
  d.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  d.location.id = device;
  d.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  rc = access(va, bytes, &d, 1);
  if (rc != CUDA_SUCCESS) return -96000 - static_cast<int>(rc);
  g_maps[static_cast<int64_t>(va)] = bytes;
  return static_cast<int64_t>(va);
}

extern "C" int prl_nvls_unmap(int64_t va) {
  auto unmap = drv<FnUnmap>("cuMemUnmap");
  auto afree = drv<FnAddrFree>("cuMemAddressFree");
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_maps.find(va);
  if (!unmap || !afree || it == g_maps.end()) return -97;
  unmap(static_cast<CUdeviceptr>(va), it->second);
  afree(static_cast<CUdeviceptr>(va), it->second);
  g_maps.erase(it);
  return 0;
}

extern "C" int prl_nvls_release(int64_t id) {
  auto rel = drv<FnRelease>("cuMemRelease");
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_handles.find(id);
  if (!rel || it == g_handles.end()) return -98;
  rel(it->second.h);
  g_handles.erase(it);
  return 0;
}

extern "C" int prl_close_fd(int fd) { return close(fd); }

// =================================================================================================
// multimem kernels
// =================================================================================================
namespace prl {

struct Bf16x8 {
  uint32_t v[4];
};
// sum over all devices of 8 bf16 values at the same multicast offset, accumulated in fp32 inside the switch
__device__ __forceinline__ Bf16x8 mm_ld_reduce_bf16x8(const void* mc) {
  Bf16x8 r;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.acc::f32.v4.bf16x2 {%0, %1, %2, %3}, [%4];"
               : "=r"(r.v[0]), "=r"(r.v[1]), "=r"(r.v[2]), "=r"(r.v[3])
               : "l"(mc)
               : "memory");
  return r;
}
__device__ __forceinline__ void mm_st_bf16x8(void* mc, const Bf16x8& r) {
  asm volatile("multimem.st.relaxed.sys.global.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(mc), "r"(r.v[0]), "r"(r.v[1]),
               "r"(r.v[2]), "r"(r.v[3])
               : "memory");
}
__device__ __forceinline__ void mm_st_v4(void* mc, const int4& v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc), "f"(__int_as_float(v.x)),
               "f"(__int_as_float(v.y)), "f"(__int_as_float(v.z)), "f"(__int_as_float(v.w))
               : "memory");
}

constexpr int kMaxRanks = 16;
struct FlagPtrs {
  uint32_t* p[kMaxRanks];
};

// all ranks of the launch meet here: rank r bumps flag[r] on every peer, then waits for every peer's bump.
// `epoch` is device resident (graph replays advance it); only block 0 talks to the fabric, the grid follows through `gate`.
__device__ __forceinline__ void grid_rank_barrier(const FlagPtrs& flags, int rank, int world, uint32_t epoch, uint32_t* gate,
                                                  uint32_t gate_target) {
  __syncthreads();
  if (blockIdx.x == 0) {
    if (threadIdx.x < world) {
      __threadfence_system();
      asm volatile("red.release.sys.global.max.u32 [%0], %1;" ::"l"(flags.p[threadIdx.x] + rank), "r"(epoch) : "memory");
      wait_ge_sys(flags.p[rank] + threadIdx.x, epoch);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      __threadfence();
      atomicMax(gate, gate_target);
    }
  }
  if (threadIdx.x == 0) {
    uint32_t spins = 0;
    while (static_cast<int32_t>(*reinterpret_cast<volatile uint32_t*>(gate) - gate_target) < 0) {
      if (++spins > PRL_SPIN_LIMIT) { printf("[prl] nvls gate timeout\n"); __trap(); }
    }
    __threadfence();
  }
  __syncthreads();
}

// The LAST CTA of the launch publishes the new epoch: every CTA has read the old one by then (state[2] = done counter).
__device__ __forceinline__ void publish_epoch(uint32_t* state, uint32_t e) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(state + 2, 1u) == gridDim.x - 1) {
      state[2] = 0;
      *reinterpret_cast<volatile uint32_t*>(state) = e;
      __threadfence();
    }
  }
}

// Two-shot NVLS all-reduce of n16 x 16 bytes of bf16 living at `mc_in` (multicast view of every rank's input buffer):
// rank r reduces the r-th slice in the switch and multicast-stores the result into every rank's output buffer
// (`mc_out`; may alias the input), optionally adding a local residual while it holds the sum.
//   state[0] = epoch counter, state[1] = grid gate (both device resident, monotone)
__global__ void __launch_bounds__(512)
nvls_allreduce_kernel(const uint8_t* __restrict__ mc_in, uint8_t* __restrict__ mc_out, const uint8_t* __restrict__ residual,
                      int64_t n16, FlagPtrs flags, int rank, int world, uint32_t* __restrict__ state) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(state) + 2;    // two barrier rounds per launch
  __syncthreads();
  const uint32_t e = s_epoch;
  grid_rank_barrier(flags, rank, world, e - 1, state + 1, e - 1);      // every rank's input is written and visible
  const int64_t per = (n16 + world - 1) / world;
  const int64_t lo = per * rank, hi = min(n16, lo + per);
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = lo + blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < hi; i += stride) {
    Bf16x8 s = mm_ld_reduce_bf16x8(mc_in + i * 16);
    if (residual != nullptr) {
      const int4 q = ld_nc_v4(residual + i * 16);
      const uint32_t* qu = reinterpret_cast<const uint32_t*>(&q);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16(s.v[j]), b = unpack_bf16(qu[j]);
        s.v[j] = pack_bf16(a.x + b.x, a.y + b.y);
      }
    }
    mm_st_bf16x8(mc_out + i * 16, s);
  }
  __threadfence_system();
  grid_rank_barrier(flags, rank, world, e, state + 1, e);              // every slice has landed everywhere
  publish_epoch(state, e);
}

// One-shot variant for latency-bound sizes (TP decode activations, <= ~2 MB): ONE rank barrier, then every rank reduces the
// WHOLE buffer in the switch and keeps the sum locally (+ optional residual), so nothing has to become visible on a peer
// afterwards.  The caller alternates two input buffers (a rank may rewrite buffer p for call c+2 only after the barrier of
// call c+1, which every peer reaches after it has finished reading p in call c).
__global__ void __launch_bounds__(512)
nvls_allreduce_oneshot_kernel(const uint8_t* __restrict__ mc_in, uint8_t* __restrict__ out, const uint8_t* __restrict__ residual,
                              int64_t n16, FlagPtrs flags, int rank, int world, uint32_t* __restrict__ state) {
  __shared__ uint32_t s_epoch;
  if (threadIdx.x == 0) s_epoch = *reinterpret_cast<volatile uint32_t*>(state) + 1;
  __syncthreads();
  const uint32_t e = s_epoch;
  grid_rank_barrier(flags, rank, world, e, state + 1, e);              // every rank's partial sums are written and visible
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  // two multimem loads (+ their residual loads) in flight per thread: the switch round trip dominates at these sizes
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16; i += 2 * stride) {
    const int64_t i2 = i + stride;
    const bool two = i2 < n16;
    Bf16x8 s0 = mm_ld_reduce_bf16x8(mc_in + i * 16), s1;
    if (two) s1 = mm_ld_reduce_bf16x8(mc_in + i2 * 16);
    if (residual != nullptr) {
      const int4 q0 = ld_nc_v4(residual + i * 16);
      int4 q1 = q0;
      if (two) q1 = ld_nc_v4(residual + i2 * 16);
      const uint32_t* u0 = reinterpret_cast<const uint32_t*>(&q0);
      const uint32_t* u1 = reinterpret_cast<const uint32_t*>(&q1);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 a = unpack_bf16(s0.v[j]), b = unpack_bf16(u0[j]);
        s0.v[j] = pack_bf16(a.x + b.x, a.y + b.y);
        if (two) {
          const float2 c = unpack_bf16(s1.v[j]), d = unpack_bf16(u1[j]);
          s1.v[j] = pack_bf16(c.x + d.x, c.y + d.y);
        }
      }
    }
    int4 o;
    o.x = s0.v[0]; o.y = s0.v[1]; o.z = s0.v[2]; o.w = s0.v[3];
    st_v4(out + i * 16, o);
    if (two) {
      o.x = s1.v[0]; o.y = s1.v[1]; o.z = s1.v[2]; o.w = s1.v[3];
      st_v4(out + i2 * 16, o);
    }
  }
  publish_epoch(state, e);
}

// NVLS reduce-scatter fused with the fp32 gradient accumulate: acc[i] += scale * sum_r grads_r[shard_off + i] for the shard this
// rank owns, the sum taken in the switch (one multimem load per 8 elements instead of world-1 peer loads).
__global__ void __launch_bounds__(512)
nvls_reduce_scatter_accum_kernel(const uint8_t* __restrict__ mc_grads, int64_t shard_off_elems, float* __restrict__ acc,
                                 int64_t n, float scale, float* __restrict__ sumsq) {
  const int64_t n8 = n >> 3;
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  float local = 0.f;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n8; i += stride) {
    const Bf16x8 s = mm_ld_reduce_bf16x8(mc_grads + (shard_off_elems + i * 8) * 2);
    float4 a0 = *reinterpret_cast<const float4*>(acc + i * 8), a1 = *reinterpret_cast<const float4*>(acc + i * 8 + 4);
    const float2 f0 = unpack_bf16(s.v[0]), f1 = unpack_bf16(s.v[1]), f2 = unpack_bf16(s.v[2]), f3 = unpack_bf16(s.v[3]);
    a0.x += scale * f0.x; a0.y += scale * f0.y; a0.z += scale * f1.x; a0.w += scale * f1.y;
    a1.x += scale * f2.x; a1.y += scale * f2.y; a1.z += scale * f3.x; a1.w += scale * f3.y;
    *reinterpret_cast<float4*>(acc + i * 8) = a0;
    *reinterpret_cast<float4*>(acc + i * 8 + 4) = a1;
    local += a0.x * a0.x + a0.y * a0.y + a0.z * a0.z + a0.w * a0.w + a1.x * a1.x + a1.y * a1.y + a1.z * a1.z + a1.w * a1.w;
  }
  if (sumsq != nullptr) {
    local = warp_sum(local);
    if ((threadIdx.x & 31) == 0) atomicAdd(sumsq, local);
  }
}

// Multicast all-gather / broadcast: store this rank's shard at its slot of EVERY rank's buffer with one instruction per 16 bytes.
__global__ void __launch_bounds__(512)
nvls_bcast_store_kernel(const int4* __restrict__ src, uint8_t* __restrict__ mc_dst, int64_t n16) {
  const int64_t stride = static_cast<int64_t>(gridDim.x) * blockDim.x;
  for (int64_t i = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; i < n16; i += stride)
    mm_st_v4(mc_dst + i * 16, ld_nc_v4(src + i));
}

}  // namespace prl

using namespace prl;

extern "C" int prl_nvls_allreduce(const void* mc_in, void* mc_out, const void* residual, int64_t bytes, void* const* flag_ptrs,
                                  int rank, int world, uint32_t* state, int ctas, cudaStream_t stream) {
  if (bytes % 16 != 0 || world > kMaxRanks) return -100;
  FlagPtrs f;
  for (int i = 0; i < kMaxRanks; ++i) f.p[i] = i < world ? static_cast<uint32_t*>(flag_ptrs[i]) : nullptr;
  const int64_t n16 = bytes / 16;
  const int64_t per = (n16 + world - 1) / world;
  int grid = ctas > 0 ? ctas : static_cast<int>((per + 511) / 512);
  if (grid < 1) grid = 1;
  if (grid > 64) grid = 64;                      // all CTAs must be co-resident: they meet at a grid gate
  nvls_allreduce_kernel<<<grid, 512, 0, stream>>>(static_cast<const uint8_t*>(mc_in), static_cast<uint8_t*>(mc_out),
                                                  static_cast<const uint8_t*>(residual), n16, f, rank, world, state);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int prl_nvls_allreduce_oneshot(const void* mc_in, void* out, const void* residual, int64_t bytes, void* const* flag_ptrs,
                                          int rank, int world, uint32_t* state, int ctas, cudaStream_t stream) {
  if (bytes % 16 != 0 || world > kMaxRanks) return -103;
  FlagPtrs f;
  for (int i = 0; i < kMaxRanks; ++i) f.p[i] = i < world ? static_cast<uint32_t*>(flag_ptrs[i]) : nullptr;
  const int64_t n16 = bytes / 16;
  int grid = ctas > 0 ? ctas : static_cast<int>((n16 + 511) / 512);
  if (grid < 1) grid = 1;
  if (grid > 64) grid = 64;                      // all CTAs must be co-resident: they meet at a grid gate
  nvls_allreduce_oneshot_kernel<<<grid, 512, 0, stream>>>(static_cast<const uint8_t*>(mc_in), static_cast<uint8_t*>(out),
                                                          static_cast<const uint8_t*>(residual), n16, f, rank, world, state);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int prl_nvls_reduce_scatter_accum(const void* mc_grads, int64_t shard_off_elems, float* acc, int64_t n, float scale,
                                             float* sumsq, int ctas, cudaStream_t stream) {
  if (n % 8 != 0 || shard_off_elems % 8 != 0) return -101;
  int grid = ctas > 0 ? ctas : 64;
  nvls_reduce_scatter_accum_kernel<<<grid, 512, 0, stream>>>(static_cast<const uint8_t*>(mc_grads), shard_off_elems, acc, n, scale,
                                                             sumsq);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}

extern "C" int prl_nvls_bcast_store(const void* src, void* mc_dst, int64_t bytes, int ctas, cudaStream_t stream) {
  if (bytes % 16 != 0) return -102;
  int grid = ctas > 0 ? ctas : 64;
  nvls_bcast_store_kernel<<<grid, 512, 0, stream>>>(static_cast<const int4*>(src), static_cast<uint8_t*>(mc_dst), bytes / 16);
  PRL_CUDA_CHECK(cudaGetLastError());
  return 0;
}
