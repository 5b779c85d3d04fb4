"""Symmetric device heap over NVLink/NVSwitch: every rank allocates the same-sized buffer, exchanges CUDA
IPC handles through the (gloo/NCCL) process group, and maps every peer's buffer, so kernels can issue
plain ld/st (or TMA) on peer HBM.  This is the substrate for the in-kernel collectives (csrc/p2p.cu,
csrc/ag_gemm.cu) that replace the reference's NCCL / TCP / shm paths.

Only ranks on one node (one NVSwitch domain) can share a heap; multi-node jobs fall back to NCCL.
"""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.distributed as dist

from ..ops import ext

_DTYPE_SIZE = {torch.bfloat16: 2, torch.float16: 2, torch.float32: 4, torch.uint8: 1, torch.int32: 4,
               torch.int64: 8, torch.float8_e4m3fn: 1}


@dataclass
class SymmBuffer:
    """One symmetric allocation: local tensor view + device pointers of the same buffer on every rank."""
    local: torch.Tensor
    ptrs: list[int]            # index = rank in the heap's group
    nbytes: int

    def peer_ptr(self, rank: int, byte_offset: int = 0) -> int:
        return self.ptrs[rank] + byte_offset

    def peer_ptrs(self, byte_offset: int = 0) -> list[int]:
        return [p + byte_offset for p in self.ptrs]


class SymmetricHeap:
    def __init__(self, group=None, device: torch.device | None = None):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._C = ext()
        self._allocs: list[tuple[int, list[int]]] = []
        self._flags = None
        self._epoch = 0
        if self.world > 1:
            devs: list = [None] * self.world
            dist.all_gather_object(devs, self.device.index, group=group)
            for d in devs:
                if d != self.device.index:
                    self._C.enable_peer_access(int(d))

    def alloc(self, numel: int, dtype: torch.dtype = torch.bfloat16) -> SymmBuffer:
        nbytes = ((numel * _DTYPE_SIZE[dtype] + 255) // 256) * 256
        with torch.cuda.device(self.device):
            ptr = self._C.sym_malloc(nbytes)
        ptrs = [ptr]
        opened: list[int] = []
        if self.world > 1:
            handle = self._C.ipc_get_handle(ptr)
            handles: list = [None] * self.world
            dist.all_gather_object(handles, handle, group=self.group)
            ptrs = []
            for r, h in enumerate(handles):
                if r == self.rank:
                    ptrs.append(ptr)
                else:
                    p = self._C.ipc_open_handle(h)
                    ptrs.append(p)
                    opened.append(p)
        self._allocs.append((ptr, opened))
        local = self._C.tensor_from_ptr(ptr, [numel], dtype, self.device.index)
        return SymmBuffer(local=local, ptrs=ptrs, nbytes=nbytes)

    def view_peer(self, buf: SymmBuffer, rank: int, numel: int, dtype: torch.dtype, elem_offset: int = 0) -> torch.Tensor:
        """Tensor view of a peer's buffer (debug / tests; kernels use raw pointers)."""
        return self._C.tensor_from_ptr(buf.ptrs[rank] + elem_offset * _DTYPE_SIZE[dtype], [numel], dtype,
                                       self.device.index)

    def barrier(self, channel: int = 0) -> None:
        """Device-side barrier across the heap's ranks on the current stream (no host sync).

        Barriers on one *channel* must all be issued on the same CUDA stream (the device-side epoch advances in execution order);
        use a different channel per stream — the ZeRO-3 engine uses channel 1 for its communication stream."""
        if self.world == 1:
            return
        if self._flags is None:
            self._flags = self.alloc(64 * 4, torch.int32)        # 4 channels x 64 flag words
            # device-resident epoch counter per channel (the barrier kernel bumps it: CUDA-graph replays stay in step)
            self._epochs = [torch.zeros(1, dtype=torch.int32, device=self.device) for _ in range(4)]
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)       # everyone's flag words are zeroed and mapped
        self._C.p2p_barrier(self._flags.peer_ptrs(channel * 64 * 4), self.rank, self._epochs[channel])

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        for ptr, opened in self._allocs:
            for p in opened:
                try:
                    self._C.ipc_close_handle(p)
                except Exception:
                    pass
        if self.world > 1 and dist.is_initialized():
            try:
                dist.barrier(group=self.group)
            except Exception:
                pass
        for ptr, _ in self._allocs:
            try:
                self._C.sym_free(ptr)
            except Exception:
                pass
        self._allocs.clear()


def make_heap(group=None, device: torch.device | None = None):
    """Symmetric heap factory: the VMM / multicast heap (``parallel/nvls.py``) where the driver supports it, the legacy CUDA-IPC
    heap otherwise or when ``POLYRL_HEAP=ipc``.  Every rank of ``group`` takes the same branch (the support flag is all-reduced)."""
    import os

    if os.environ.get("POLYRL_HEAP", "nvls") != "ipc":
        try:
            from .nvls import NvlsHeap

            ok = torch.tensor([1 if ext().nvls_supported((device or torch.device("cuda", torch.cuda.current_device())).index) else 0],
                              device=device, dtype=torch.int32)
            if dist.is_initialized() and (dist.get_world_size(group) > 1):
                dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
            if int(ok.item()):
                return NvlsHeap(group=group, device=device)
        except Exception as e:      # noqa: BLE001
            print(f"[symm] NVLS heap unavailable ({e}); using the CUDA-IPC heap", flush=True)
    return SymmetricHeap(group=group, device=device)
