"""Native in-process rollout engine: paged KV cache, continuous batching, CUDA-graph decode.

Replaces the patched SGLang server the reference launches per TP group
(rlboost/verl_stream/workers/rollout/sglang_rollout/sglang_rollout_remote.py:103-182,
rlboost/verl_stream/workers/rollout/sglang_rollout/sglang_http_async_engine.py:43-113).  The engine
lives in the trainer process, so colocated weight hand-off is pointer aliasing (no CUDA-IPC
serialisation, SURVEY §2.5c last rows) and ``release/resume_memory_occupation`` is dropping /
re-creating the KV pool.

Per step the scheduler either *prefills* newly admitted prompts (packed varlen forward that also fills
the KV pages) or *decodes* one token for every running sequence.  The decode step — embedding, all
layers (fused GEMM epilogues, RoPE+KV append, paged flash-decode), LM head, Gumbel-max sampling with
log-prob, and the on-device bookkeeping for the next step — is captured once per batch-size bucket in a
CUDA graph and replayed; the host only feeds page-table deltas and reads back tokens one step late.
"""
from __future__ import annotations

import itertools
import os
import time
import zlib
from collections import deque
from dataclasses import dataclass, field

import torch

from .. import ops
from ..models import transformer as tf
from ..models.config import ModelConfig

PAGE = 16


@dataclass
class SamplingParams:
    max_new_tokens: int = 128
    temperature: float = 1.0
    top_p: float = 1.0
    top_k: int = -1
    n: int = 1
    stop_token_ids: tuple = ()
    ignore_eos: bool = False
    seed: int | None = None

    @classmethod
    def from_dict(cls, d: dict | None) -> "SamplingParams":
        d = dict(d or {})
        keep = {k: d[k] for k in ("max_new_tokens", "temperature", "top_p", "top_k", "n", "ignore_eos", "seed") if k in d}
        if "stop_token_ids" in d and d["stop_token_ids"]:
            keep["stop_token_ids"] = tuple(d["stop_token_ids"])
        return cls(**keep)


@dataclass
class Request:
    rid: str
    prompt_ids: list[int]
    params: SamplingParams
    output_ids: list[int] = field(default_factory=list)
    logprobs: list[float] = field(default_factory=list)
    finish_reason: str | None = None       # "stop" | "length" | "abort"
    slot: int = -1
    blocks: list[int] = field(default_factory=list)
    shared_prefix_blocks: int = 0
    t_arrival: float = 0.0
    group: object = None                   # requests of one n-sample group share prompt KV pages
    donor: object = None                   # the group member whose prefill this request re-uses
    cached_tokens: int = 0                 # prompt tokens whose KV came from the prefix cache (multiple of the page size)
    prefilled: int = -1                    # prompt tokens whose KV is in the cache (-1: not admitted yet); < len(prompt) while
                                           # a long prompt is being prefilled chunk by chunk
    # vision-language prompts: {"pixel_values": [N, patch_dim], "grid_thw": [(t, h, w), ...]}; the prompt carries one placeholder
    # token per merged patch.  mrope_pos [3, T] / mrope_delta come from models.vision.mrope_positions at admission.
    mm: dict | None = None
    mrope_pos: object = None
    mrope_delta: int = 0
    image_embeds: object = None            # tower output, computed at the first prefill chunk and dropped after the last
    h: int = -1                            # handle inside the native scheduler (csrc/scheduler.hpp), -1 = not registered

    @property
    def decoding(self) -> bool:
        return self.prefilled >= len(self.prompt_ids) and self.finish_reason is None

    @property
    def finished(self) -> bool:
        return self.finish_reason is not None

    @property
    def num_tokens(self) -> int:
        return len(self.prompt_ids) + len(self.output_ids)


class BlockAllocator:
    """Ref-counted KV page allocator; block 0 is a scratch page for idle batch slots."""

    def __init__(self, num_blocks: int):
        self.num_blocks = num_blocks
        self.free = deque(range(1, num_blocks))
        self.ref = [0] * num_blocks

    def available(self) -> int:
        return len(self.free)

    def alloc(self, n: int) -> list[int]:
        if n > len(self.free):
            raise MemoryError("KV cache exhausted")
        out = [self.free.popleft() for _ in range(n)]
        for b in out:
            self.ref[b] = 1
        return out

    def share(self, blocks: list[int]) -> None:
        for b in blocks:
            self.ref[b] += 1

    def release(self, blocks: list[int]) -> None:
        for b in blocks:
            self.ref[b] -= 1
            if self.ref[b] == 0:
                self.free.append(b)


class PrefixCache:
    """Cross-request KV reuse at page granularity (the role of SGLang's radix cache for the reference's multi-turn /
    shared-system-prompt traffic).  A full page is identified by the hash chain of every token up to and including it, so a hit
    on page i implies hits on pages 0..i-1.  The cache holds one reference on every registered block; blocks nobody else
    references are evicted in LRU order when the allocator runs dry.  Must be flushed when the weights change."""

    def __init__(self, alloc: BlockAllocator):
        self.alloc = alloc
        self.by_key: dict[int, int] = {}
        self.key_of: dict[int, int] = {}
        self.lru: "dict[int, None]" = {}            # insertion-ordered: oldest first
        self.hits = self.misses = 0

    @staticmethod
    def page_keys(tokens: list[int], max_pages: int) -> list[int]:
        keys, h = [], 0
        for i in range(max_pages):
            h = hash((h, tuple(tokens[i * PAGE:(i + 1) * PAGE])))
            keys.append(h)
        return keys

    def match(self, keys: list[int]) -> list[int]:
        out = []
        for k in keys:
            b = self.by_key.get(k)
            if b is None:
                break
            out.append(b)
        for b in out:                               # refresh recency
            self.lru.pop(b, None)
            self.lru[b] = None
        self.hits += len(out)
        self.misses += len(keys) - len(out)
        return out

    def insert(self, keys: list[int], blocks: list[int]) -> None:
        for k, b in zip(keys, blocks):
            if k in self.by_key or b in self.key_of:
                continue
            self.by_key[k], self.key_of[b] = b, k
            self.alloc.share([b])
            self.lru[b] = None

    def evict(self, want: int) -> int:
        freed = 0
        for b in list(self.lru):
            if freed >= want:
                break
            if self.alloc.ref[b] == 1:              # only the cache holds it
                self._drop(b)
                freed += 1
        return freed

    def _drop(self, b: int) -> None:
        self.by_key.pop(self.key_of.pop(b), None)
        self.lru.pop(b, None)
        self.alloc.release([b])

    def flush(self) -> None:
        for b in list(self.key_of):
            self._drop(b)


class NativeBlockAllocator:
    """``BlockAllocator`` API over the C++ ``KvPageManager`` (csrc/kv_manager.hpp)."""

    def __init__(self, num_blocks: int, mgr=None):
        from ..ops import ext

        self.num_blocks = num_blocks
        self.mgr = mgr or ext().KvPageManager(num_blocks, PAGE)

    def available(self) -> int:
        return self.mgr.available()

    def alloc(self, n: int) -> list[int]:
        try:
            return self.mgr.alloc(n)
        except RuntimeError as e:
            raise MemoryError(str(e)) from None

    def share(self, blocks: list[int]) -> None:
        self.mgr.share(blocks)

    def release(self, blocks: list[int]) -> None:
        self.mgr.release(blocks)


class NativePrefixCache:
    """``PrefixCache`` API over the same C++ object as the allocator (page keys are hashed natively, one call per request)."""

    def __init__(self, alloc: NativeBlockAllocator):
        self.alloc, self.mgr = alloc, alloc.mgr

    def page_keys(self, tokens: list[int], max_pages: int) -> list[int]:
        return self.mgr.page_keys(tokens, max_pages)

    def match(self, keys: list[int]) -> list[int]:
        return self.mgr.match(keys)

    def insert(self, keys: list[int], blocks: list[int]) -> None:
        self.mgr.insert(keys, blocks)

    def evict(self, want: int) -> int:
        return self.mgr.evict(want)

    def flush(self) -> None:
        self.mgr.flush()

    @property
    def by_key(self):
        return range(self.mgr.cached_pages())

    @property
    def hits(self) -> int:
        return self.mgr.hits

    @property
    def misses(self) -> int:
        return self.mgr.misses


def make_kv_manager(num_blocks: int, prefix_cache: bool):
    """(allocator, prefix cache or None): the native page manager unless ``POLYRL_NATIVE_KV=0`` or the extension is not built."""
    if os.environ.get("POLYRL_NATIVE_KV", "1") == "1":
        try:
            alloc = NativeBlockAllocator(num_blocks)
            return alloc, (NativePrefixCache(alloc) if prefix_cache else None)
        except Exception:            # noqa: BLE001  (no extension in this environment: pure-Python bookkeeping)
            if torch.cuda.is_available():
                raise
    alloc = BlockAllocator(num_blocks)
    return alloc, (PrefixCache(alloc) if prefix_cache else None)


class RolloutEngine:
    def __init__(self, cfg: ModelConfig, device, max_num_seqs: int = 256, max_model_len: int = 2048,
                 kv_cache_tokens: int | None = None, gpu_memory_utilization: float = 0.5, dtype=None,
                 use_cuda_graph: bool = True, max_prefill_tokens: int = 16384, seed: int = 0, tp=None,
                 weight_dtype: str = "bf16", steps_per_sync: int | None = None, enable_prefix_cache: bool = False,
                 kv_vmm: bool | None = None):
        if cfg.sliding_window and max_model_len > cfg.sliding_window:
            raise NotImplementedError(f"max_model_len {max_model_len} exceeds the model's attention window {cfg.sliding_window}: "
                                      "windowed decode attention is not implemented (full-causal only)")
        # decode steps launched back to back per host round trip (tokens are read back once per window; a sequence that
        # ends mid-window rides along for a few ignored tokens).  1 = one-step lag only.
        import os as _os
        self.steps_per_sync = max(1, min(8, int(steps_per_sync or _os.environ.get("POLYRL_DECODE_STEPS_PER_SYNC", "1"))))
        # KV pool in a CUDA-VMM range: release_memory / resume_memory swap the PHYSICAL pages only, addresses (and therefore
        # captured decode graphs) stay valid across PPO steps (utils/vmm.py)
        import os as _os2
        self.kv_vmm = (kv_vmm if kv_vmm is not None else _os2.environ.get("POLYRL_KV_VMM", "0") == "1") and str(device).startswith("cuda")
        self._vmm = None
        self._kv_views: tuple[list, list] | None = None
        self._kv_blocks = 0
        self.enable_prefix_cache = enable_prefix_cache and tp is None and cfg.family == "llama"
        self.prefix_cache: PrefixCache | None = None
        if weight_dtype not in ("bf16", "fp8"):
            raise ValueError("weight_dtype must be 'bf16' or 'fp8'")
        if weight_dtype == "fp8" and tp is not None:
            raise NotImplementedError("fp8 rollout weights with tensor parallelism")
        self.weight_dtype = weight_dtype               # fp8: projections re-quantised (e4m3, per-row scale) at every sync
        self._fp8_cache: list[dict] = []
        self.tp = tp                               # rollout.tp.TPContext for tensor-parallel engines
        self.full_cfg = cfg
        cfg = tp.local_cfg() if tp is not None else cfg   # per-rank view: local heads / kv heads / intermediate
        self.cfg = cfg
        self.device = torch.device(device)
        self.is_cuda = self.device.type == "cuda"
        self.dtype = dtype or (torch.bfloat16 if self.is_cuda else torch.float32)
        self.max_num_seqs = max_num_seqs
        self.max_model_len = max_model_len
        self.max_blocks_per_seq = (max_model_len + PAGE - 1) // PAGE
        self.max_prefill_tokens = max_prefill_tokens
        # MoE blocks group tokens by expert with data-dependent shapes (host-side counts): their decode step cannot be captured
        # (POLYRL_MOE_STATIC_DECODE=1 switches decode to the static-shape expert path, which can)
        self.use_cuda_graph = use_cuda_graph and self.is_cuda and (not cfg.num_experts or tf._MOE_STATIC_DECODE())
        self.gpu_memory_utilization = gpu_memory_utilization
        self._kv_tokens_req = kv_cache_tokens
        self.embed_w: dict | None = None
        self.layer_w: list[dict] | None = None
        self.head_w: dict | None = None
        self.weight_version = 0
        self.k_cache: list[torch.Tensor] = []
        self.v_cache: list[torch.Tensor] = []
        self.alloc: BlockAllocator | None = None
        self.sched = None                              # native scheduler (csrc/scheduler.hpp), created with the KV pool
        self.waiting: deque[Request] = deque()
        self.running: dict[int, Request] = {}          # slot -> request
        self.free_slots = list(range(max_num_seqs - 1, -1, -1))
        self._rid = itertools.count()
        self._graphs: dict[int, tuple] = {}
        self._pending = None        # (event, cpu tokens, cpu logprobs, slots snapshot) of the step in flight
        self._prefill_queue: list[Request] = []     # admitted requests whose prompt is not completely prefilled yet
        self._prefill_turn = True                    # alternate prefill chunks with decode steps
        self._base_seed = seed
        self._step_count = 0
        self._top_k, self._top_p = -1, 1.0
        self.vision_w = None                         # VLM tower weights (set_vision_weights)
        self.stats = {"prefill_tokens": 0, "decode_tokens": 0, "steps": 0, "last_gen_throughput": 0.0}
        self._alloc_static()

    # ------------------------------------------------------------------ weights / memory
    def set_weights(self, embed_w: dict, layer_w: list[dict], head_w: dict, version: int | None = None,
                    presharded: bool = False) -> None:
        """Point the engine at (possibly trainer-owned) weight tensors.  Pointers that change invalidate
        captured graphs; in-place updates of the same storage do not."""
        old = self._weight_ptrs()
        if self.tp is not None and not presharded:
            layer_w = [self.tp.shard_layer(w) for w in layer_w]
        if self.weight_dtype == "fp8":
            layer_w = self._quantize_layers(layer_w)
        self.embed_w, self.layer_w, self.head_w = embed_w, layer_w, head_w
        if old != self._weight_ptrs():
            self._graphs.clear()
        self.flush_cache()              # same storage updated in place still means new weights
        if version is not None:
            self.weight_version = version

    _FP8_KEYS = ("wqkv", "wo", "wgu", "wd")

    @torch.no_grad()
    def _quantize_layers(self, layer_w: list[dict]) -> list[dict]:
        """bf16 -> (e4m3, row scale) for the four block projections; buffers are reused across syncs so captured decode
        graphs stay valid (reference capability: fp8 rollout with bf16 update, BASELINE config 4)."""
        fresh = len(self._fp8_cache) != len(layer_w)
        if fresh:
            self._fp8_cache = [dict() for _ in layer_w]
        out = []
        for w, cache in zip(layer_w, self._fp8_cache):
            nw = dict(w)
            for k in self._FP8_KEYS:
                if k not in w:
                    continue
                src = w[k] if w[k].dtype in (torch.bfloat16, torch.float32) else w[k].float()
                fw = ops.quantize_weight_fp8(src.to(self.dtype) if src.dtype != self.dtype else src)
                if k in cache and cache[k].q.shape == fw.q.shape:
                    cache[k].q.copy_(fw.q)
                    cache[k].scale.copy_(fw.scale)
                else:
                    cache[k] = fw
                nw[k] = cache[k]
            out.append(nw)
        return out

    def _weight_ptrs(self):
        if self.layer_w is None:
            return None
        return tuple((t.q if isinstance(t, ops.Fp8Weight) else t).data_ptr()
                     for w in [self.embed_w, *self.layer_w, self.head_w] for t in w.values())

    def kv_bytes_per_token(self) -> int:
        c = self.cfg
        heads, width, two = c.kv_cache_layout()          # MLA: one shared latent entry per token, no separate value pool
        return (2 if two else 1) * c.num_layers * heads * width * torch.empty(0, dtype=self.dtype).element_size()

    def init_kv_cache(self) -> None:
        """(Re)create the paged KV pool (``resume_memory_occupation(['kv_cache'])``)."""
        if self.k_cache:
            return
        if self.is_cuda:
            torch.cuda.empty_cache()       # hand the trainer's cached activation blocks back before sizing the pool
        if self._kv_tokens_req is not None:
            tokens = self._kv_tokens_req
        elif self.is_cuda:
            free, total = torch.cuda.mem_get_info(self.device)
            tokens = int(free * self.gpu_memory_utilization) // self.kv_bytes_per_token()
        else:
            tokens = self.max_num_seqs * self.max_model_len
        tokens = min(tokens, self.max_num_seqs * self.max_blocks_per_seq * PAGE + PAGE)
        num_blocks = max(2, tokens // PAGE)
        c = self.cfg
        kv_heads, kv_width, two_pools = c.kv_cache_layout()
        shape = (num_blocks, kv_heads, PAGE, kv_width)
        if not two_pools:
            # latent attention: the value pool is a prefix view of the key pool (models/mla.py), nothing extra is allocated
            self.k_cache = [torch.zeros(shape, device=self.device, dtype=self.dtype) for _ in range(c.num_layers)]
            self.v_cache = [k[..., : c.kv_lora_rank] for k in self.k_cache]
            self.alloc, self.prefix_cache = make_kv_manager(num_blocks, self.enable_prefix_cache)
            self._make_scheduler()
            self._graphs.clear()
            return
        if self.kv_vmm:
            if self._vmm is None:                      # sized once; later resumes re-back the same virtual range
                from ..utils.vmm import VmmBuffer

                per = num_blocks * c.num_kv_heads * PAGE * c.head_dim * torch.empty(0, dtype=self.dtype).element_size()
                per = (per + 255) // 256 * 256
                self._vmm = VmmBuffer(2 * c.num_layers * per, self.device)
                self._vmm.map()
                self._kv_views = ([self._vmm.tensor((2 * i) * per, shape, self.dtype) for i in range(c.num_layers)],
                                  [self._vmm.tensor((2 * i + 1) * per, shape, self.dtype) for i in range(c.num_layers)])
                self._kv_blocks = num_blocks
                self._graphs.clear()
            else:
                self._vmm.map()
            self.k_cache, self.v_cache = self._kv_views
            for t in (*self.k_cache, *self.v_cache):
                t.zero_()                              # fresh physical pages: masked tail slots must hold finite values
            self.alloc, self.prefix_cache = make_kv_manager(self._kv_blocks, self.enable_prefix_cache)
            self._make_scheduler()
            return
        self.k_cache = [torch.zeros(shape, device=self.device, dtype=self.dtype) for _ in range(c.num_layers)]
        self.v_cache = [torch.zeros(shape, device=self.device, dtype=self.dtype) for _ in range(c.num_layers)]
        self.alloc, self.prefix_cache = make_kv_manager(num_blocks, self.enable_prefix_cache)
        self._make_scheduler()
        self._graphs.clear()

    def _make_scheduler(self) -> None:
        """Native admission scheduler over the native page manager (csrc/scheduler.hpp); ``POLYRL_NATIVE_SCHED=0`` keeps the Python
        reference implementation below (`_admit`, the planning half of `_prefill`, `_finish`), which the tests compare it against."""
        self.sched = None
        if os.environ.get("POLYRL_NATIVE_SCHED", "1") == "1" and isinstance(self.alloc, NativeBlockAllocator):
            if self.running:
                raise RuntimeError("KV pool re-created with sequences running")
            self.sched = ops.ext().EngineScheduler(self.alloc.mgr, self.max_num_seqs, PAGE, self.prefix_cache is not None)
            self.free_slots = list(range(self.max_num_seqs - 1, -1, -1))
            for r in self.waiting:
                r.h = -1

    def release_memory(self, tags=("kv_cache",)) -> None:
        """Free engine memory so the trainer can use it (colocated mode).  Requires no running requests."""
        if "kv_cache" in tags:
            if self.running or self.waiting:
                raise RuntimeError("cannot release the KV cache with requests in flight")
            self.k_cache, self.v_cache, self.alloc, self.prefix_cache, self.sched = [], [], None, None, None
            if self.kv_vmm and self._vmm is not None:
                torch.cuda.synchronize(self.device)        # nothing may still read the pool
                self._vmm.unmap()                          # graphs stay valid: same addresses after the next resume
            else:
                self._graphs.clear()
            if self.is_cuda:
                torch.cuda.empty_cache()
        if "weights" in tags:
            self.embed_w = self.layer_w = self.head_w = None
            self._graphs.clear()

    def resume_memory(self, tags=("kv_cache",)) -> None:
        if "kv_cache" in tags:
            self.init_kv_cache()

    def flush_cache(self) -> None:
        """Drop every cached prefix page (mandatory after a weight update: the KV was produced by the old weights)."""
        if self.prefix_cache is not None:
            self.prefix_cache.flush()

    # ------------------------------------------------------------------ static decode state
    def _alloc_static(self) -> None:
        s, dev = self.max_num_seqs, self.device
        self.st_input_ids = torch.zeros(s, dtype=torch.long, device=dev)
        self.st_positions = torch.zeros(s, dtype=torch.long, device=dev)
        self.st_seq_lens = torch.zeros(s, dtype=torch.int32, device=dev)
        self.st_active = torch.zeros(s, dtype=torch.long, device=dev)
        self.st_block_table = torch.zeros(s, self.max_blocks_per_seq + 2, dtype=torch.int32, device=dev)   # + overshoot slack -> scratch page 0
        self.st_temp = torch.ones(s, dtype=torch.float32, device=dev)
        self.st_seeds = torch.zeros(s, dtype=torch.long, device=dev)
        self.st_out_tok = torch.zeros(s, dtype=torch.long, device=dev)
        self.st_out_lp = torch.zeros(s, dtype=torch.float32, device=dev)
        pin = self.is_cuda
        k = self.steps_per_sync
        self.h_tok = [torch.zeros(k, s, dtype=torch.long, pin_memory=pin) for _ in range(2)]
        self.h_lp = [torch.zeros(k, s, dtype=torch.float32, pin_memory=pin) for _ in range(2)]

    # ------------------------------------------------------------------ request API
    def set_vision_weights(self, vision_w) -> None:
        """Tower weights (models.vision.VisionWeights-like: ``.stem`` dict + ``.blocks`` list) for image prompts."""
        self.vision_w = vision_w

    def add_request(self, prompt_ids: list[int], params: SamplingParams | dict | None = None, rid: str | None = None,
                    group=None, mm: dict | None = None) -> Request:
        if not isinstance(params, SamplingParams):
            params = SamplingParams.from_dict(params)
        mrope_pos, mrope_delta = None, 0
        if mm is not None:
            from ..models import vision as V

            if self.full_cfg.vision is None:
                raise ValueError(f"model {self.full_cfg.name!r} has no vision tower; image inputs are not accepted")
            if len(prompt_ids) + 1 > self.max_model_len:
                raise ValueError("image prompt longer than max_model_len (image token runs cannot be truncated)")
            prompt_ids = list(prompt_ids)
            n_ph = sum(t in (self.full_cfg.image_token_id, self.full_cfg.video_token_id) for t in prompt_ids)
            if n_ph == len(mm["grid_thw"]):                        # one placeholder per image / video: expand to one per merged patch
                prompt_ids = V.expand_image_placeholders(self.full_cfg, prompt_ids, mm["grid_thw"])
            mrope_pos, mrope_delta = V.mrope_positions(self.full_cfg, prompt_ids, mm["grid_thw"], mm.get("second_per_grid_ts"))
        if len(prompt_ids) + 1 > self.max_model_len:
            prompt_ids = prompt_ids[-(self.max_model_len - 1):]
        budget = self.max_model_len - len(prompt_ids)
        if params.max_new_tokens > budget:
            params = SamplingParams(**{**params.__dict__, "max_new_tokens": budget})
        r = Request(rid or f"r{next(self._rid)}", list(prompt_ids), params, t_arrival=time.time(), group=group, mm=mm,
                    mrope_pos=mrope_pos, mrope_delta=mrope_delta)
        self.waiting.append(r)
        return r

    def abort(self, rid: str | None = None, abort_all: bool = False) -> list[Request]:
        out = []
        for r in list(self.waiting):
            if abort_all or r.rid == rid:
                self.waiting.remove(r)
                if r.h >= 0 and self.sched is not None:
                    self.sched.remove_waiting(r.h)
                r.h = -1
                r.finish_reason = "abort"
                out.append(r)
        self._drain_pending()
        for slot, r in list(self.running.items()):
            if abort_all or r.rid == rid:
                self._finish(r, "abort")
                out.append(r)
        return out

    def has_work(self) -> bool:
        return bool(self.waiting or self.running or self._pending or self._prefill_queue)

    def num_running(self) -> int:
        return len(self.running)

    def num_waiting(self) -> int:
        return len(self.waiting)

    # ------------------------------------------------------------------ scheduling
    def _blocks_needed(self, r: Request) -> int:
        total = len(r.prompt_ids) + r.params.max_new_tokens
        return (total + PAGE - 1) // PAGE

    def _admit(self) -> list[Request]:
        """Pop waiting requests that fit (slots, KV pages reserved up front, prefill token budget)."""
        if self.sched is not None:
            return self._admit_native()
        admitted, tokens = [], 0
        shared: dict = {}
        while self.waiting and self.free_slots:
            r = self.waiting[0]
            need = self._blocks_needed(r)
            full_prompt_pages = len(r.prompt_ids) // PAGE
            donor = shared.get(id(r.group)) if r.group is not None else None
            cached: list[int] = []
            keys: list[int] = []
            if donor is None and self.prefix_cache is not None and r.mm is None:     # placeholder ids do not identify an image
                # at least the last prompt token must be recomputed (its logits seed the first sampled token)
                keys = self.prefix_cache.page_keys(r.prompt_ids, (len(r.prompt_ids) - 1) // PAGE)
                cached = self.prefix_cache.match(keys)
                self.alloc.share(cached)            # pin before any eviction below can touch them
            fresh = need - (full_prompt_pages if donor is not None else len(cached))
            if fresh > self.alloc.available() and self.prefix_cache is not None:
                self.prefix_cache.evict(fresh - self.alloc.available())
            if fresh > self.alloc.available():
                if cached:
                    self.alloc.release(cached)
                break
            if tokens >= self.max_prefill_tokens and admitted:      # this round's chunk budget is spent
                if cached:
                    self.alloc.release(cached)
                break
            self.waiting.popleft()
            if donor is not None:
                pre = donor.blocks[:full_prompt_pages]
                self.alloc.share(pre)
                r.blocks = pre + self.alloc.alloc(fresh)
                r.shared_prefix_blocks = full_prompt_pages
                r.donor = donor
                r.prefilled = 0
            else:
                r.blocks = cached + self.alloc.alloc(fresh)
                r.cached_tokens = r.prefilled = len(cached) * PAGE
                if r.group is not None:
                    shared[id(r.group)] = r
            r.slot = self.free_slots.pop()
            self.running[r.slot] = r
            admitted.append(r)
            tokens += 0 if donor is not None else len(r.prompt_ids) - r.prefilled
        return admitted

    def _admit_native(self) -> list[Request]:
        """``_admit`` through the C++ scheduler: same decisions, the per-request results are mirrored onto the Request objects."""
        for r in self.waiting:                     # registration is lazy so the native queue always has the order of self.waiting
            if r.h < 0:
                r.h = self.sched.add(r.prompt_ids, r.params.max_new_tokens, id(r.group) if r.group is not None else -1, r.mm is None)
        admitted, by_h = [], {}
        for h in self.sched.admit(self.max_prefill_tokens):
            r = self.waiting.popleft()
            slot, blocks, shared, cached, donor = self.sched.info(h)
            if r.h != h or self.free_slots.pop() != slot:
                raise RuntimeError("native scheduler out of step with the engine's request queue")
            r.slot, r.blocks, r.shared_prefix_blocks = slot, blocks, shared
            if donor >= 0:
                r.donor, r.prefilled = by_h[donor], 0
            else:
                r.cached_tokens = r.prefilled = cached
            by_h[h] = r
            self.running[slot] = r
            admitted.append(r)
        return admitted

    def _finish(self, r: Request, reason: str) -> None:
        r.finish_reason = reason
        if r.slot >= 0:
            self.running.pop(r.slot, None)
            self.st_active[r.slot] = 0
            self.st_seq_lens[r.slot] = 0
            self.free_slots.append(r.slot)
            r.slot = -1
        if r.h >= 0 and self.sched is not None:
            to_cache = self.prefix_cache is not None and reason != "abort" and r.mm is None
            self.sched.finish(r.h, (r.prompt_ids + r.output_ids) if to_cache else [], to_cache)
            r.h, r.blocks = -1, []
            return
        if r.blocks and self.alloc is not None:
            if self.prefix_cache is not None and reason != "abort" and r.mm is None:
                # every FULL page whose KV is final: prompt + generated tokens except the last one (never fed back)
                toks = r.prompt_ids + r.output_ids
                n_full = max(0, (len(toks) - 1)) // PAGE
                self.prefix_cache.insert(self.prefix_cache.page_keys(toks, n_full), r.blocks[:n_full])
            self.alloc.release(r.blocks)
        r.blocks = []

    # ------------------------------------------------------------------ forward passes
    def _lm_logits(self, h: torch.Tensor) -> torch.Tensor:
        w = tf.lm_head_weight(self.cfg, self.head_w, self.embed_w)
        return ops.linear(h, w.to(h.dtype))

    @torch.no_grad()
    def _prefill(self, reqs: list[Request]) -> None:
        cfg, dev = self.cfg, self.device
        # n-sample groups: only the donor runs the prompt; followers share its full KV pages, copy its partial last
        # page and sample their own first token from the donor's last-position logits.
        # Chunked prefill: at most ``max_prefill_tokens`` new tokens per call; a prompt longer than the remaining budget is
        # cut, the rest is processed by later calls (attending to the KV already in the cache) interleaved with decode
        # steps.  Requests whose prompt is complete after this call get their first token sampled below.
        self._prefill_queue = [r for r in self._prefill_queue if not r.finished]
        for r in reqs:                          # a follower whose donor was aborted mid-prefill becomes its own donor
            if r.donor is not None and r.donor.finished and r.donor.prefilled < len(r.donor.prompt_ids):
                r.donor, r.prefilled = None, 0
        budget = self.max_prefill_tokens if self.tp is None else 1 << 60      # the TP prefill path does not chunk
        donors, spans = [], []
        for r in reqs:
            if r.donor is not None or r.finished:
                continue
            c = r.prefilled
            e = min(len(r.prompt_ids), c + max(budget, 0))
            if e <= c:
                continue                        # no budget left for this one in this round
            donors.append(r)
            spans.append((c, e))
            budget -= e - c
        if not donors:
            return
        d_index = {id(r): i for i, r in enumerate(donors)}
        ids, pos, slots, lens = [], [], [], []
        any_cached = any(c > 0 for c, _ in spans)
        g_blk, g_off, klens = [], [], []
        native = self.sched is not None and all(r.h >= 0 for r in donors)
        if native:       # one pass in C++ over every token of the chunk (csrc/scheduler.hpp::prefill_arrays)
            ids_t, pos_t, slots_t, lens_t, g_blk, g_off, klens_t = self.sched.prefill_arrays(
                [(r.h, c, e) for r, (c, e) in zip(donors, spans)], any_cached)
            lens, klens = lens_t.tolist(), klens_t.tolist()
            for r, (c, e) in zip(donors, spans):
                r.prefilled = e
        else:
            for r, (c, e) in zip(donors, spans):
                ids += r.prompt_ids[c:e]
                pos += list(range(c, e))
                slots += [r.blocks[t // PAGE] * PAGE + t % PAGE for t in range(c, e)]
                lens.append(e - c)
                if any_cached:
                    g_blk += [r.blocks[t // PAGE] for t in range(e)]
                    g_off += [t % PAGE for t in range(e)]
                    klens.append(e)
                r.prefilled = e
            ids_t, pos_t, slots_t = (torch.tensor(v, dtype=torch.long) for v in (ids, pos, slots))
            if any_cached:
                g_blk, g_off = torch.tensor(g_blk, dtype=torch.long), torch.tensor(g_off, dtype=torch.long)
        cu = torch.zeros(len(donors) + 1, dtype=torch.int32)
        cu[1:] = torch.tensor(lens, dtype=torch.int32).cumsum(0)
        inp = tf.PackedInput(ids_t.to(dev), pos_t.to(dev), cu.to(dev), max(lens))
        if any(r.mm is not None for r in donors):
            inp.mrope_pos, inp.image_embeds = self._mm_prefill_inputs(donors, spans)
        if any_cached:
            cu_k = torch.zeros(len(donors) + 1, dtype=torch.int32)
            cu_k[1:] = torch.tensor(klens, dtype=torch.int32).cumsum(0)
            inp.kv_gather = (g_blk.to(dev), g_off.to(dev), cu_k.to(dev), max(klens))
        slot_t = slots_t.to(dev)
        x = tf.embed_forward(cfg, self.embed_w, inp).to(self.dtype)
        for i, w in enumerate(self.layer_w):
            if self.tp is not None:
                from .tp import layer_prefill_tp

                x = layer_prefill_tp(self.tp, cfg, w, x, inp, (self.k_cache[i], self.v_cache[i], slot_t))
            else:
                x = tf.layer_forward(cfg, w, x, inp, kv_sink=(self.k_cache[i], self.v_cache[i], slot_t))
            if inp.image_embeds is not None:
                from ..models.vision import deepstack_add

                x = deepstack_add(self.full_cfg, x, inp.input_ids, inp.image_embeds, i)      # Qwen3-VL DeepStack (no-op otherwise)
        self.stats["prefill_tokens"] += int(ids_t.numel())
        done_ids = {id(r) for r in donors if r.prefilled >= len(r.prompt_ids)}
        reqs = [r for r in reqs if not r.finished and (id(r) in done_ids or (r.donor is not None and id(r.donor) in done_ids))]
        self._prefill_queue = [r for r in self._prefill_queue if not r.finished and not any(r is q for q in reqs)]
        if not reqs:
            return
        last = (cu[1:] - 1).long().to(dev)
        h = tf.final_norm(cfg, self.head_w, x.index_select(0, last))
        logits_d = self._lm_logits(h)
        src_blk, dst_blk = [], []
        for r in reqs:
            if r.donor is not None:
                r.prefilled = len(r.prompt_ids)
            if r.donor is not None and len(r.prompt_ids) % PAGE != 0:
                fp = len(r.prompt_ids) // PAGE
                src_blk.append(r.donor.blocks[fp])
                dst_blk.append(r.blocks[fp])
        if dst_blk:
            s_t = torch.tensor(src_blk, dtype=torch.long, device=dev)
            d_t = torch.tensor(dst_blk, dtype=torch.long, device=dev)
            for i in range(len(self.layer_w)):
                self.k_cache[i].index_copy_(0, d_t, self.k_cache[i].index_select(0, s_t))
                self.v_cache[i].index_copy_(0, d_t, self.v_cache[i].index_select(0, s_t))
        row = torch.tensor([d_index[id(r.donor if r.donor is not None else r)] for r in reqs], dtype=torch.long, device=dev)
        logits = logits_d.index_select(0, row)
        temps = torch.tensor([r.params.temperature for r in reqs], dtype=torch.float32, device=dev)
        seeds = torch.tensor([self._seed_for(r, 0) for r in reqs], dtype=torch.long, device=dev)
        p0 = reqs[0].params
        tok, lp = ops.sample_tokens(logits, temps, seeds, greedy=False, top_k=p0.top_k, top_p=p0.top_p)
        tok_c, lp_c = tok.tolist(), lp.tolist()
        for r in reqs:
            r.donor = None            # release the reference; pages stay shared through the allocator's ref-counts
        live, rows = [], []
        for r, t, l in zip(reqs, tok_c, lp_c):
            self._append_token(r, t, l)
            if r.finished:
                continue
            n = r.num_tokens            # tokens so far; the new token sits at position n-1 (+ the M-RoPE offset of image prompts)
            live.append((r.slot, t, n - 1 + r.mrope_delta, n, 1, r.params.temperature, self._seed_for(r, 1)))
            rows.append(r)
        if live:
            cols = list(zip(*live))
            sl = torch.tensor(cols[0], dtype=torch.long, device=dev)
            self.st_input_ids[sl] = torch.tensor(cols[1], dtype=torch.long, device=dev)
            self.st_positions[sl] = torch.tensor(cols[2], dtype=torch.long, device=dev)
            self.st_seq_lens[sl] = torch.tensor(cols[3], dtype=torch.int32, device=dev)
            self.st_active[sl] = torch.tensor(cols[4], dtype=torch.long, device=dev)
            self.st_temp[sl] = torch.tensor(cols[5], dtype=torch.float32, device=dev)
            self.st_seeds[sl] = torch.tensor(cols[6], dtype=torch.long, device=dev)
            width = self.st_block_table.shape[1]
            if self.sched is not None and all(r.h >= 0 for r in rows):
                table = self.sched.block_rows([r.h for r in rows], width)
            else:
                table = torch.tensor([r.blocks + [0] * (width - len(r.blocks)) for r in rows], dtype=torch.int32)
            self.st_block_table[sl] = table.to(dev)

    def _mm_prefill_inputs(self, donors: list[Request], spans: list[tuple[int, int]]):
        """M-RoPE positions [3, T] and tower embeddings for the image-token rows of this prefill chunk.  The tower runs once per
        request (first chunk) over all of its images; later chunks take their slice of the cached rows."""
        from ..models import vision as V

        cfg, dev = self.full_cfg, self.device
        if getattr(self, "vision_w", None) is None:
            raise RuntimeError("image request but no vision weights were set (RolloutEngine.set_vision_weights)")
        pos, emb = [], []
        for r, (c, e) in zip(donors, spans):
            if r.mm is None:
                pos.append(torch.arange(c, e).expand(3, -1))
                continue
            pos.append(r.mrope_pos[:, c:e])
            if r.image_embeds is None:
                pv = r.mm["pixel_values"].to(dev, self.dtype)
                with torch.no_grad():
                    r.image_embeds = V.vision_forward(cfg, self.vision_w, pv, r.mm["grid_thw"])
            ph = [t in (cfg.image_token_id, cfg.video_token_id) for t in r.prompt_ids]
            lo, hi = sum(ph[:c]), sum(ph[:e])
            if hi > lo:
                emb.append(r.image_embeds[lo:hi])
            if e >= len(r.prompt_ids):
                r.image_embeds = None
        pos_t = torch.cat(pos, dim=1).contiguous().to(dev)
        emb_t = torch.cat(emb, dim=0) if emb else torch.zeros(0, V.embed_width(cfg), device=dev, dtype=self.dtype)
        return pos_t, emb_t

    def _seed_for(self, r: Request, step: int) -> int:
        base = r.params.seed if r.params.seed is not None else self._base_seed
        return (zlib.crc32(f"{base}/{r.rid}".encode()) << 12) + step * 1000003

    def lockstep_digest(self) -> int:
        """Order-sensitive checksum of every token this engine has appended so far: the ranks of one TP group must agree on it
        (bench.py selfcheck, tests/tp_worker.py)."""
        return int(getattr(self, "_tok_digest", 0))

    def _append_token(self, r: Request, tok: int, lp: float) -> None:
        self._tok_digest = (getattr(self, "_tok_digest", 0) * 1000003 + int(tok) + 1) & 0x7FFFFFFFFFFFFFFF
        r.output_ids.append(int(tok))
        r.logprobs.append(float(lp))
        p = r.params
        if not p.ignore_eos and (tok == self.cfg.eos_token_id or tok in p.stop_token_ids or tok in self.cfg.extra_eos_token_ids):
            self._finish(r, "stop")
        elif len(r.output_ids) >= p.max_new_tokens or r.num_tokens >= self.max_model_len:
            self._finish(r, "length")

    def _decode_body(self, bs: int) -> None:
        """One decode step over the first ``bs`` batch slots (captured into a CUDA graph)."""
        cfg = self.cfg
        ids, pos = self.st_input_ids[:bs], self.st_positions[:bs]
        bt, sl, act = self.st_block_table[:bs], self.st_seq_lens[:bs], self.st_active[:bs]
        # KV index of the new token: equals its rotary position for text models; image prompts shift the rotary position by the
        # request's M-RoPE delta, so vision-language engines derive the index from the sequence length instead
        kvpos = pos if self.full_cfg.vision is None else (sl.long() - 1).clamp_(min=0)
        page_idx = torch.div(kvpos, PAGE, rounding_mode="floor").clamp_(max=self.max_blocks_per_seq)
        blk = bt.gather(1, page_idx.unsqueeze(1)).squeeze(1).long()
        slot = (blk * PAGE + kvpos % PAGE) * act          # idle slots write into scratch page 0
        inp = tf.DecodeInput(ids, pos, slot, bt, sl, self.max_model_len)
        x = torch.nn.functional.embedding(ids, self.embed_w["embed"]).to(self.dtype)
        if cfg.family == "gpt2":
            x = x + torch.nn.functional.embedding(pos, self.embed_w["pos_embed"]).to(self.dtype)
        fused = (self.tp is None and cfg.family == "llama" and self.is_cuda and self.dtype == torch.bfloat16 and not cfg.num_experts
                 and cfg.attn_type == "mha" and cfg.mlp_act == "swiglu")
        delta = None
        for i, w in enumerate(self.layer_w):
            if self.tp is not None:
                from .tp import layer_decode_tp

                x = layer_decode_tp(self.tp, cfg, w, x, inp, self.k_cache[i], self.v_cache[i])
            elif fused:
                x, delta = tf.layer_decode_fused(cfg, w, x, delta, inp, self.k_cache[i], self.v_cache[i])
            else:
                x = tf.layer_decode(cfg, w, x, inp, self.k_cache[i], self.v_cache[i])
        if delta is not None:
            h = ops.ext().add_rmsnorm_(x, delta, self.head_w["norm_w"], cfg.rms_norm_eps)
        else:
            h = tf.final_norm(cfg, self.head_w, x)
        logits = self._lm_logits(h)
        tok, lp = ops.sample_tokens(logits, self.st_temp[:bs], self.st_seeds[:bs], top_k=self._top_k, top_p=self._top_p)
        self.st_out_tok[:bs].copy_(tok)
        self.st_out_lp[:bs].copy_(lp)
        # on-device bookkeeping for the next step
        self.st_input_ids[:bs].copy_(tok)
        self.st_positions[:bs].add_(act)
        self.st_seq_lens[:bs].add_(act.to(torch.int32))
        self.st_seeds[:bs].add_(1000003)

    def _bucket(self, n: int) -> int:
        b = 8
        while b < n:
            b *= 2
        return min(b, self.max_num_seqs)

    def _run_decode(self, bs: int) -> None:
        if not self.use_cuda_graph or self._top_k > 0 or self._top_p < 1.0:
            # top-k / top-p go through the sort-based library sampler (host-seeded RNG): not captured into a graph
            self._decode_body(bs)
            return
        key = (bs, self._top_k, self._top_p)
        g = self._graphs.get(key)
        if g is None:
            # warm-up on a side stream (also sets kernel attributes), then capture
            snap = [t.clone() for t in (self.st_input_ids, self.st_positions, self.st_seq_lens, self.st_seeds)]
            s = torch.cuda.Stream(device=self.device)
            s.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(s):
                self._decode_body(bs)
            torch.cuda.current_stream(self.device).wait_stream(s)
            for dst, src in zip((self.st_input_ids, self.st_positions, self.st_seq_lens, self.st_seeds), snap):
                dst.copy_(src)
            graph = torch.cuda.CUDAGraph()
            n0 = ops.total_launches()
            with torch.cuda.graph(graph):
                self._decode_body(bs)
            per_replay = ops.total_launches() - n0
            for dst, src in zip((self.st_input_ids, self.st_positions, self.st_seq_lens, self.st_seeds), snap):
                dst.copy_(src)
            self._graphs[key] = (graph, per_replay)
            g = self._graphs[key]
        g[0].replay()
        ops.LAUNCHES["graph_replayed_kernels"] = ops.LAUNCHES.get("graph_replayed_kernels", 0) + g[1]

    def _compact_slots(self) -> None:
        """Keep running sequences in the lowest slots so a smaller graph bucket can be used."""
        if not self.running:
            return
        n = len(self.running)
        bucket = self._bucket(n)
        high = sorted(s for s in self.running if s >= bucket)
        if not high:
            return
        low_free = sorted(s for s in range(bucket) if s not in self.running)
        if self.sched is not None:
            self.sched.move_slots([(self.running[src].h, dst) for src, dst in zip(high, low_free)])
        for src, dst in zip(high, low_free):
            r = self.running.pop(src)
            for t in (self.st_input_ids, self.st_positions, self.st_seq_lens, self.st_active, self.st_temp,
                      self.st_seeds):
                t[dst] = t[src]
            self.st_block_table[dst] = self.st_block_table[src]
            self.st_active[src] = 0
            self.st_seq_lens[src] = 0
            r.slot = dst
            self.running[dst] = r
            self.free_slots.remove(dst)
            self.free_slots.append(src)
        self.free_slots.sort(reverse=True)

    def _drain_pending(self) -> list[Request]:
        """Consume the tokens of the decode step in flight; returns requests that finished."""
        done = []
        if self._pending is None:
            return done
        ev, h_tok, h_lp, slots = self._pending
        self._pending = None
        if ev is not None:
            t0 = time.perf_counter()
            ev.synchronize()
            self.stats["gpu_wait_s"] = self.stats.get("gpu_wait_s", 0.0) + time.perf_counter() - t0     # ~0 => host-bound loop
        toks, lps = h_tok.tolist(), h_lp.tolist()      # one bulk conversion instead of 2 tensor-scalar reads per request
        for s, r in slots:
            if r.finished:
                continue
            for tk, lp in zip(toks, lps):                # the window's steps, oldest first
                self._append_token(r, tk[s], lp[s])
                if r.finished:
                    done.append(r)
                    break
        return done

    @torch.no_grad()
    def step(self) -> list[Request]:
        """Run one scheduler iteration; returns requests that finished during it."""
        if self.layer_w is None:
            raise RuntimeError("RolloutEngine has no weights (call set_weights)")
        self.init_kv_cache()
        finished: list[Request] = []
        decoding = any(r.decoding for r in self.running.values())
        if self._prefill_queue and (self._prefill_turn or not decoding):
            # continue a chunked prefill (alternating with decode steps so running sequences keep advancing)
            finished = self._drain_pending()
            self._prefill_turn = False
            batch = list(self._prefill_queue)
            self._prefill(batch)
            finished += [r for r in batch if r.finished]
            self.stats["steps"] += 1
            return finished
        self._prefill_turn = True
        if not self._prefill_queue and self.waiting and self.free_slots and self.alloc.available() > 0:
            finished = self._drain_pending()
            admitted = self._admit()
            if admitted:
                self._top_k, self._top_p = admitted[0].params.top_k, admitted[0].params.top_p
                self._prefill_queue = list(admitted)
                self._prefill(admitted)
                finished += [r for r in admitted if r.finished]
                self.stats["steps"] += 1
                return finished
        if not decoding:
            return finished + self._drain_pending()
        self._compact_slots()
        bs = self._bucket(max(self.running) + 1)
        # launch step k+1 first, then consume step k's tokens while the GPU works (one-step lag:
        # a sequence that ended at step k rides along for one harmless extra token)
        i = self._step_count & 1
        self._step_count += 1
        for j in range(self.steps_per_sync):
            self._run_decode(bs)
            self.h_tok[i][j].copy_(self.st_out_tok, non_blocking=True)
            self.h_lp[i][j].copy_(self.st_out_lp, non_blocking=True)
        ev = None
        if self.is_cuda:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
        live = [(s_, r_) for s_, r_ in self.running.items() if r_.decoding]      # mid-prefill requests hold idle slots
        new_pending = (ev, self.h_tok[i], self.h_lp[i], live)
        self.stats["decode_tokens"] += len(live) * self.steps_per_sync
        finished += self._drain_pending()
        self._pending = new_pending
        self.stats["steps"] += 1
        return finished

    # ------------------------------------------------------------------ batch API
    @torch.no_grad()
    def generate(self, prompts: list[list[int]], params: SamplingParams | dict | None = None,
                 on_finish=None, mm: list[dict | None] | None = None) -> list[list[Request]]:
        """Generate ``params.n`` samples for every prompt; returns, per prompt, its list of requests.  ``mm[i]`` carries the image
        inputs of prompt i (vision-language models)."""
        if not isinstance(params, SamplingParams):
            params = SamplingParams.from_dict(params)
        groups: list[list[Request]] = []
        for pi, p in enumerate(prompts):
            tag = object()
            m = mm[pi] if mm is not None else None
            groups.append([self.add_request(p, params, group=tag if params.n > 1 else None, mm=m) for _ in range(params.n)])
        t0 = time.time()
        ntok0 = self.stats["decode_tokens"]
        while self.has_work():
            for r in self.step():
                if on_finish is not None:
                    on_finish(r)
        dt = max(time.time() - t0, 1e-9)
        self.stats["last_gen_throughput"] = (self.stats["decode_tokens"] - ntok0) / dt
        return groups

    def server_info(self) -> dict:
        return {"#running_req": len(self.running), "#queue_req": len(self.waiting),
                "last_gen_throughput": self.stats["last_gen_throughput"], "weight_version": self.weight_version}
