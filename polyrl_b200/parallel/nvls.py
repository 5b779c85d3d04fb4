"""NVLS symmetric heap: CUDA-VMM allocations shared between the ranks of one NVSwitch domain as POSIX file descriptors,
mapped three ways on every rank —

* ``local``      this rank's share (a normal tensor),
* ``ptrs[r]``    every peer's share, for plain ld/st / TMA over NVLink (what ``SymmetricHeap`` offered through legacy CUDA IPC),
* ``mc_ptr``     ONE multicast address that means "this offset on every GPU": ``multimem.ld_reduce`` sums the N copies inside
                 the switch, ``multimem.st`` writes all N copies (csrc/nvls.cu).

File descriptors travel over ``AF_UNIX`` sockets with ``SCM_RIGHTS`` (``socket.send_fds``); fabric handles are not needed on one
node.  If the driver / container has no multicast support the heap still works with ``mc_ptr = 0`` and callers fall back to
the peer-pointer kernels.

Reference: NVLS is reachable in the reference only through NCCL inside SGLang (``NCCL_NVLS_ENABLE``,
3rdparty/verl/verl/workers/rollout/sglang_rollout/sglang_rollout.py:99); SURVEY §5.8 asks for the in-kernel form.
"""
from __future__ import annotations

import os
import socket
import tempfile
import time
import uuid
from dataclasses import dataclass, field

import torch
import torch.distributed as dist

from ..ops import ext
from .symm import _DTYPE_SIZE, SymmBuffer


@dataclass
class NvlsBuffer(SymmBuffer):
    mc_ptr: int = 0                       # multicast VA (0 = multicast unavailable)
    handles: list = field(default_factory=list)
    maps: list = field(default_factory=list)

    def mc(self, byte_offset: int = 0) -> int:
        return self.mc_ptr + byte_offset


def _send_fd_to_all(path: str, fd: int, n_clients: int) -> None:
    srv = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
    if os.path.exists(path):
        os.unlink(path)
    srv.bind(path)
    srv.listen(n_clients)
    try:
        for _ in range(n_clients):
            c, _ = srv.accept()
            socket.send_fds(c, [b"fd"], [fd])
            c.close()
    finally:
        srv.close()
        os.unlink(path)


def _recv_fd(path: str, timeout: float = 60.0) -> int:
    t0 = time.time()
    while True:
        try:
            c = socket.socket(socket.AF_UNIX, socket.SOCK_STREAM)
            c.connect(path)
            break
        except (FileNotFoundError, ConnectionRefusedError):
            c.close()
            if time.time() - t0 > timeout:
                raise TimeoutError(f"no fd server at {path}")
            time.sleep(0.01)
    _, fds, _, _ = socket.recv_fds(c, 16, 1)
    c.close()
    return fds[0]


class NvlsHeap:
    """Drop-in superset of ``SymmetricHeap`` (same ``alloc`` / ``barrier`` / ``close``) whose buffers also carry ``mc_ptr``."""

    def __init__(self, group=None, device: torch.device | None = None, multicast: bool = True):
        self.group = group
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.device = device or torch.device("cuda", torch.cuda.current_device())
        self._C = ext()
        self._bufs: list[NvlsBuffer] = []
        self._flags = None
        self._epochs = None
        tag = [uuid.uuid4().hex[:12]]
        if self.world > 1:
            dist.broadcast_object_list(tag, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        self._tag = tag[0]
        self._seq = 0
        want_mc = multicast and self.world > 1 and bool(self._C.nvls_supported(self.device.index))
        if self.world > 1:           # every rank must agree, or the collective setup below deadlocks
            flag = torch.tensor([1 if want_mc else 0], device=self.device, dtype=torch.int32)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=group)
            want_mc = bool(flag.item())
            devs: list = [None] * self.world
            dist.all_gather_object(devs, self.device.index, group=group)
            self._devs = devs
            for d in devs:
                if d != self.device.index:
                    self._C.enable_peer_access(int(d))
        else:
            self._devs = [self.device.index]
        self.multicast = want_mc
        self.multicast_error: str | None = None

    def _sock(self, kind: str, src: int) -> str:
        return os.path.join(tempfile.gettempdir(), f"prl_nvls_{self._tag}_{self._seq}_{kind}_{src}.sock")

    def _exchange(self, kind: str, my_fd: int | None, srcs: list[int]) -> dict[int, int]:
        """Every rank in ``srcs`` serves its fd to all other ranks; returns {src: received fd}."""
        got: dict[int, int] = {}
        for s in srcs:
            if s == self.rank:
                _send_fd_to_all(self._sock(kind, s), my_fd, self.world - 1)
            else:
                got[s] = _recv_fd(self._sock(kind, s))
        return got

    def alloc(self, numel: int, dtype: torch.dtype = torch.bfloat16, multicast: bool = True) -> NvlsBuffer:
        C, dev = self._C, self.device.index
        nbytes = numel * _DTYPE_SIZE[dtype]
        size = int(C.nvls_round_size(max(nbytes, 1), dev, self.world))
        if size <= 0:
            raise RuntimeError(f"nvls_round_size failed ({size})")
        self._seq += 1
        mem = int(C.nvls_create(size, dev))
        if mem <= 0:
            raise RuntimeError(f"cuMemCreate failed ({mem})")
        handles, maps = [mem], []
        local_va = int(C.nvls_map(mem, dev))
        if local_va <= 0:
            raise RuntimeError(f"cuMemMap failed ({local_va})")
        maps.append(local_va)
        ptrs = [local_va]
        mc_va = 0
        if self.world > 1:
            fd = int(C.nvls_export_fd(mem))
            if fd < 0:
                raise RuntimeError(f"cuMemExportToShareableHandle failed ({fd})")
            got = self._exchange("mem", fd, list(range(self.world)))
            C.close_fd(fd)
            ptrs = []
            for r in range(self.world):
                if r == self.rank:
                    ptrs.append(local_va)
                    continue
                h = int(C.nvls_import_fd(got[r], size, dev))
                C.close_fd(got[r])
                if h <= 0:
                    raise RuntimeError(f"cuMemImportFromShareableHandle failed ({h})")
                va = int(C.nvls_map(h, dev))
                if va <= 0:
                    raise RuntimeError(f"cuMemMap(peer) failed ({va})")
                handles.append(h)
                maps.append(va)
                ptrs.append(va)
            if self.multicast and multicast:
                mc_va = self._setup_multicast(size, mem, handles, maps)
        local = C.tensor_from_ptr(local_va, [numel], dtype, dev)
        local.zero_()
        torch.cuda.synchronize(self.device)
        if self.world > 1:
            dist.barrier(group=self.group)
        buf = NvlsBuffer(local=local, ptrs=ptrs, nbytes=size, mc_ptr=mc_va, handles=handles, maps=maps)
        self._bufs.append(buf)
        return buf

    def _setup_multicast(self, size: int, mem: int, handles: list, maps: list) -> int:
        C, dev = self._C, self.device.index
        err = torch.zeros(1, device=self.device, dtype=torch.int32)
        mc = 0
        try:
            if self.rank == 0:
                mc = int(C.nvls_mc_create(size, self.world))
                fd = int(C.nvls_export_fd(mc)) if mc > 0 else -1
                ok = [mc > 0 and fd >= 0]
            else:
                ok = [True]
            # rank 0 tells everyone whether the object exists before anyone waits on a socket
            dist.broadcast_object_list(ok, src=dist.get_global_rank(self.group, 0) if self.group is not None else 0, group=self.group)
            if not ok[0]:
                raise RuntimeError(f"cuMulticastCreate failed ({mc})")
            if self.rank == 0:
                _send_fd_to_all(self._sock("mc", 0), fd, self.world - 1)
                C.close_fd(fd)
            else:
                got = _recv_fd(self._sock("mc", 0))
                mc = int(C.nvls_import_fd(got, size, dev))
                C.close_fd(got)
                if mc <= 0:
                    raise RuntimeError(f"import multicast handle failed ({mc})")
            rc = int(C.nvls_mc_add_device(mc, dev))
            if rc != 0:
                raise RuntimeError(f"cuMulticastAddDevice failed ({rc})")
        except Exception as e:          # noqa: BLE001 - reported, then agreed on collectively below
            self.multicast_error = str(e)
            err.fill_(1)
        dist.all_reduce(err, op=dist.ReduceOp.MAX, group=self.group)          # also the "every device added" barrier
        if int(err.item()):
            self.multicast = False
            return 0
        rc = int(C.nvls_mc_bind(mc, mem))
        va = int(C.nvls_map(mc, dev)) if rc == 0 else 0
        err.fill_(0 if (rc == 0 and va > 0) else 1)
        dist.all_reduce(err, op=dist.ReduceOp.MAX, group=self.group)
        if int(err.item()):
            self.multicast_error = f"cuMulticastBindMem/map failed (rc={rc}, va={va})"
            self.multicast = False
            return 0
        handles.append(mc)
        maps.append(va)
        return va

    # ---- SymmetricHeap-compatible helpers --------------------------------------------------------------------------
    def view_peer(self, buf: SymmBuffer, rank: int, numel: int, dtype: torch.dtype, elem_offset: int = 0) -> torch.Tensor:
        """Tensor view of a peer's share (debug / tests; kernels use raw pointers).  A VMM mapping reports the device that owns the
        physical memory, so the view is labelled with the PEER's device; copy it (``.to(self.device)``) before computing on it."""
        return self._C.tensor_from_ptr(buf.ptrs[rank] + elem_offset * _DTYPE_SIZE[dtype], [numel], dtype, int(self._devs[rank]))

    def barrier(self, channel: int = 0) -> None:
        if self.world == 1:
            return
        if self._flags is None:
            self._flags = self.alloc(64 * 4, torch.int32)
            self._epochs = [torch.zeros(1, dtype=torch.int32, device=self.device) for _ in range(4)]
            torch.cuda.synchronize(self.device)
            dist.barrier(group=self.group)
        self._C.p2p_barrier(self._flags.peer_ptrs(channel * 64 * 4), self.rank, self._epochs[channel])

    def close(self) -> None:
        torch.cuda.synchronize(self.device)
        if self.world > 1 and dist.is_initialized():
            try:
                dist.barrier(group=self.group)
            except Exception:
                pass
        for b in self._bufs:
            for va in b.maps:
                self._C.nvls_unmap(va)
            for h in b.handles:
                self._C.nvls_release(h)
        self._bufs.clear()


class NvlsAllReduce:
    """One-shot in-switch all-reduce of a fixed-capacity bf16 activation buffer (TP decode: o-proj / down-proj partial sums).
    ``inp`` is this rank's symmetric input buffer; the sum lands in every rank's ``out`` buffer (``multimem.st``).  CUDA-graph
    safe: epochs and the grid gate are device resident."""

    def __init__(self, heap: NvlsHeap, capacity_elems: int):
        if not heap.multicast:
            raise RuntimeError("NvlsAllReduce needs a multicast-capable heap")
        self.heap = heap
        self.inp = heap.alloc(capacity_elems, torch.bfloat16)
        self.out = heap.alloc(capacity_elems, torch.bfloat16)
        self.flags = heap.alloc(64, torch.int32)
        self.state = torch.zeros(4, dtype=torch.int32, device=heap.device)       # [epoch, grid gate, done counter, pad]
        if not (self.inp.mc_ptr and self.out.mc_ptr):
            raise RuntimeError(f"multicast mapping unavailable: {heap.multicast_error}")

    def __call__(self, numel: int, residual: torch.Tensor | None = None, ctas: int = 0) -> torch.Tensor:
        """Reduce the first ``numel`` elements of ``inp`` across ranks; returns the view of ``out`` holding the result."""
        nbytes = (numel * 2 + 15) // 16 * 16
        self.heap._C.nvls_allreduce(self.inp.mc_ptr, self.out.mc_ptr, residual, nbytes, self.flags.peer_ptrs(), self.heap.rank,
                                    self.state, ctas)
        return self.out.local[:numel]


class NvlsOneShotAllReduce:
    """Latency-optimised TP all-reduce: the producer GEMM writes its partial sums into ``next_input(numel)`` (two symmetric
    buffers alternate), then ONE kernel runs the rank barrier and reduces the whole buffer inside the switch into a local
    output (+ fused residual).  A captured CUDA graph must contain an even number of calls (decode: 2 per layer)."""

    def __init__(self, heap: NvlsHeap, capacity_elems: int):
        if not heap.multicast:
            raise RuntimeError("NvlsOneShotAllReduce needs a multicast-capable heap")
        self.heap = heap
        self.bufs = [heap.alloc(capacity_elems, torch.bfloat16) for _ in range(2)]
        self.flags = heap.alloc(64, torch.int32)
        self.state = torch.zeros(4, dtype=torch.int32, device=heap.device)
        self.turn = 0
        if not all(b.mc_ptr for b in self.bufs):
            raise RuntimeError(f"multicast mapping unavailable: {heap.multicast_error}")

    def next_input(self, numel: int) -> torch.Tensor:
        return self.bufs[self.turn].local[:numel]

    def reduce(self, out: torch.Tensor, residual: torch.Tensor | None = None, ctas: int = 0) -> torch.Tensor:
        buf = self.bufs[self.turn]
        self.turn ^= 1
        return self.heap._C.nvls_allreduce_oneshot(buf.mc_ptr, out, residual, self.flags.peer_ptrs(), self.heap.rank, self.state, ctas)
