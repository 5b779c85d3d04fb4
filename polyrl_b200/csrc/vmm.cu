// polyrl_b200 — virtual-memory-managed device buffers (CUDA VMM): a fixed virtual address range whose PHYSICAL backing can be
// released and re-created without changing any pointer.
//
// Why: in colocated RL the rollout engine's KV pool must give its HBM to the trainer every step and take it back for the next
// rollout.  Freeing and re-allocating tensors changes their addresses, which invalidates every captured decode graph (re-capture
// per PPO step) — with VMM the engine's tensors, block tables and CUDA graphs stay valid across release/resume.  This is the
// capability the reference gets from SGLang's `enable_memory_saver` / torch_memory_saver
// (rlboost/verl_stream/workers/rollout/sglang_rollout/sglang_rollout_remote.py:103-182 release/resume_memory_occupation(tags)).
//
// Driver entry points are resolved through the runtime (no link-time dependency on libcuda in the build container).
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <mutex>
#include <vector>

namespace {

template <typename Fn>
Fn drv(const char* name) {
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &p, cudaEnableDefault, &q) != cudaSuccess || !p) return nullptr;
  return reinterpret_cast<Fn>(p);
}

struct Range {
  size_t bytes = 0;
  int device = 0;
  std::vector<CUmemGenericAllocationHandle> handles;   // one per mapped chunk (empty = released)
  size_t chunk = 0;
};
std::map<CUdeviceptr, Range> g_ranges;
std::mutex g_mu;

using FnGran = CUresult (*)(size_t*, const CUmemAllocationProp*, CUmemAllocationGranularity_flags);
using FnReserve = CUresult (*)(CUdeviceptr*, size_t, size_t, CUdeviceptr, unsigned long long);
using FnAddrFree = CUresult (*)(CUdeviceptr, size_t);
using FnCreate = CUresult (*)(CUmemGenericAllocationHandle*, size_t, const CUmemAllocationProp*, unsigned long long);
using FnRelease = CUresult (*)(CUmemGenericAllocationHandle);
using FnMap = CUresult (*)(CUdeviceptr, size_t, size_t, CUmemGenericAllocationHandle, unsigned long long);
using FnUnmap = CUresult (*)(CUdeviceptr, size_t);
using FnAccess = CUresult (*)(CUdeviceptr, size_t, const CUmemAccessDesc*, size_t);

CUmemAllocationProp props(int device) {
  CUmemAllocationProp p{};
  p.type = CU_MEM_ALLOCATION_TYPE_PINNED;
  p.location.type = CU_MEM_LOCATION_TYPE_DEVICE;
  p.location.id = device;
  return p;
}

}  // namespace

extern "C" int64_t prl_vmm_granularity(int device) {
  auto f = drv<FnGran>("cuMemGetAllocationGranularity");
  if (!f) return -1;
  size_t g = 0;
  CUmemAllocationProp p = props(device);
  if (f(&g, &p, CU_MEM_ALLOC_GRANULARITY_RECOMMENDED) != CUDA_SUCCESS) return -2;
  return static_cast<int64_t>(g);
}

// Reserve `bytes` of virtual address space (rounded up to the allocation granularity).  Nothing is backed yet.
extern "C" int prl_vmm_reserve(int64_t bytes, int device, int64_t* out_ptr, int64_t* out_bytes) {
  cudaSetDevice(device);
  cudaFree(nullptr);                                        // make sure the primary context exists
  auto reserve = drv<FnReserve>("cuMemAddressReserve");
  const int64_t gran = prl_vmm_granularity(device);
  if (!reserve || gran <= 0) return -70;
  const size_t size = (static_cast<size_t>(bytes) + gran - 1) / gran * gran;
  CUdeviceptr va = 0;
  if (reserve(&va, size, static_cast<size_t>(gran), 0, 0) != CUDA_SUCCESS) return -71;
  std::lock_guard<std::mutex> lk(g_mu);
  Range r;
  r.bytes = size;
  r.device = device;
  r.chunk = static_cast<size_t>(gran) * 64;                 // map in chunks (128 MiB with 2 MiB granularity): partial failure is recoverable
  if (r.chunk > size) r.chunk = size;
  g_ranges[va] = r;
  *out_ptr = static_cast<int64_t>(va);
  *out_bytes = static_cast<int64_t>(size);
  return 0;
}

// Back the whole range with fresh physical memory (no-op when already mapped).
extern "C" int prl_vmm_map(int64_t ptr) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_ranges.find(static_cast<CUdeviceptr>(ptr));
  if (it == g_ranges.end()) return -72;
  Range& r = it->second;
  if (!r.handles.empty()) return 0;
  auto create = drv<FnCreate>("cuMemCreate");
  auto map = drv<FnMap>("cuMemMap");
  auto access = drv<FnAccess>("cuMemSetAccess");
  auto unmap = drv<FnUnmap>("cuMemUnmap");
  auto release = drv<FnRelease>("cuMemRelease");
  if (!create || !map || !access || !unmap || !release) return -73;
  CUmemAllocationProp p = props(r.device);
  CUmemAccessDesc acc{};
  acc.location = p.location;
  acc.flags = CU_MEM_ACCESS_FLAGS_PROT_READWRITE;
  size_t off = 0;
  int rc = 0;
  while (off < r.bytes) {
    const size_t n = (r.bytes - off < r.chunk) ? (r.bytes - off) : r.chunk;
    CUmemGenericAllocationHandle h;
    if (create(&h, n, &p, 0) != CUDA_SUCCESS) { rc = -74; break; }      // out of memory: roll back below
    if (map(it->first + off, n, 0, h, 0) != CUDA_SUCCESS) { release(h); rc = -75; break; }
    r.handles.push_back(h);
    off += n;
  }
  if (rc == 0 && access(it->first, r.bytes, &acc, 1) != CUDA_SUCCESS) rc = -76;
  if (rc != 0) {
    size_t o = 0;
    for (auto h : r.handles) {
      const size_t n = (r.bytes - o < r.chunk) ? (r.bytes - o) : r.chunk;
      unmap(it->first + o, n);
      release(h);
      o += n;
    }
    r.handles.clear();
  }
  return rc;
}

// Give the physical memory back; the virtual range (and every pointer into it) stays reserved.
extern "C" int prl_vmm_unmap(int64_t ptr) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_ranges.find(static_cast<CUdeviceptr>(ptr));
  if (it == g_ranges.end()) return -72;
  Range& r = it->second;
  if (r.handles.empty()) return 0;
  auto unmap = drv<FnUnmap>("cuMemUnmap");
  auto release = drv<FnRelease>("cuMemRelease");
  if (!unmap || !release) return -73;
  size_t off = 0;
  for (auto h : r.handles) {
    const size_t n = (r.bytes - off < r.chunk) ? (r.bytes - off) : r.chunk;
    unmap(it->first + off, n);
    release(h);
    off += n;
  }
  r.handles.clear();
  return 0;
}

extern "C" int prl_vmm_free(int64_t ptr) {
  prl_vmm_unmap(ptr);
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_ranges.find(static_cast<CUdeviceptr>(ptr));
  if (it == g_ranges.end()) return -72;
  auto afree = drv<FnAddrFree>("cuMemAddressFree");
  if (afree) afree(it->first, it->second.bytes);
  g_ranges.erase(it);
  return 0;
}

extern "C" int prl_vmm_is_mapped(int64_t ptr) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_ranges.find(static_cast<CUdeviceptr>(ptr));
  return it != g_ranges.end() && !it->second.handles.empty();
}
