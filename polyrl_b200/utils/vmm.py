"""Device buffers with a FIXED virtual address whose physical memory can be released and re-created (CUDA VMM, csrc/vmm.cu).

The rollout engine puts its paged KV pool in one ``VmmBuffer``: ``release_memory`` unmaps the physical pages (the trainer gets the
HBM), ``resume_memory`` maps fresh ones at the same addresses, so tensors, block tables and captured CUDA graphs survive the
round trip.  Reference capability: SGLang ``enable_memory_saver`` / torch_memory_saver release+resume of tagged regions
(rlboost/verl_stream/workers/rollout/sglang_rollout/sglang_rollout_remote.py:103-182, sglang_http_async_engine.py:257-299).
"""
from __future__ import annotations

import math

import torch

from ..ops import ext


class VmmBuffer:
    def __init__(self, nbytes: int, device):
        self.device = torch.device(device)
        self._C = ext()
        idx = self.device.index if self.device.index is not None else torch.cuda.current_device()
        self.ptr, self.nbytes = self._C.vmm_reserve(int(nbytes), idx)
        self._index = idx

    @property
    def mapped(self) -> bool:
        return bool(self.ptr) and self._C.vmm_is_mapped(self.ptr)

    def map(self) -> None:
        self._C.vmm_map(self.ptr)

    def unmap(self) -> None:
        """Callers must make sure no kernel that touches the range is still running (synchronise first)."""
        self._C.vmm_unmap(self.ptr)

    def tensor(self, byte_offset: int, shape, dtype: torch.dtype) -> torch.Tensor:
        n = math.prod(shape) * torch.empty(0, dtype=dtype).element_size()
        if byte_offset + n > self.nbytes:
            raise ValueError("view exceeds the reserved range")
        return self._C.tensor_from_ptr(self.ptr + byte_offset, list(shape), dtype, self._index)

    def free(self) -> None:
        if self.ptr:
            self._C.vmm_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass
