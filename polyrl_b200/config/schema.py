"""Typed configuration tree with the reference's key names, YAML loading and Hydra-style
dotted overrides (``actor_rollout_ref.actor.optim.lr=1e-6``).  Hydra/OmegaConf are not available on the
target image, so composition is implemented here: defaults from dataclasses < YAML file(s) < CLI
overrides, then ``validate_config`` cross-field checks (reference: 3rdparty/verl/verl/utils/config.py:23-200,
verl/trainer/config/ppo_trainer.yaml, rlboost/verl_stream/trainer/config/ppo_stream_trainer.yaml:9-23,
rlboost/verl_stream/workers/config/rollout.py:94-208).
"""
from __future__ import annotations

import copy
import os
import dataclasses
from dataclasses import dataclass, field, fields, is_dataclass
from typing import Any, Mapping

import yaml
from pathlib import Path


class BaseConfig(Mapping):
    """Dataclass mixin: mapping access (cfg['a'], cfg.get('a')), frozen-by-convention with an explicit
    allow-list of fields that may be mutated after construction (``_mutable_fields``)."""

    _mutable_fields: set = set()

    def __getitem__(self, k):
        try:
            return getattr(self, k)
        except AttributeError as e:
            raise KeyError(k) from e

    def __iter__(self):
        return (f.name for f in fields(self))

    def __len__(self):
        return len(fields(self))

    def get(self, k, default=None):
        return getattr(self, k, default)

    def to_dict(self) -> dict:
        return dataclasses.asdict(self)


@dataclass
class OptimConfig(BaseConfig):
    lr: float = 1e-6
    lr_warmup_steps: int = -1
    lr_warmup_steps_ratio: float = 0.0
    min_lr_ratio: float = 0.0
    num_cycles: float = 0.5
    warmup_style: str = "constant"      # constant | cosine
    total_training_steps: int = -1
    weight_decay: float = 0.01
    betas: tuple = (0.9, 0.999)
    eps: float = 1e-8
    clip_grad: float = 1.0
    lr_decay_style: str | None = None          # Megatron spelling of warmup_style (constant | cosine | linear): mapped in __post_init__
    lr_scheduler_type: str | None = None       # HF spelling of the same choice
    lr_warmup_init: float = 0.0
    grad_clip: float | None = None             # alias of clip_grad
    optimizer: str = "AdamW"                   # the fused AdamW kernel is the only optimizer
    optimizer_impl: str = "polyrl_b200"
    weight_decay_incr_style: str = "constant"
    lr_wsd_decay_style: str = "exponential"
    lr_wsd_decay_steps: int | None = None
    use_checkpoint_opt_param_scheduler: bool = False
    lr_decay_steps: int | None = None          # Megatron: length of the decay phase (sets total_training_steps when that is unset)
    min_lr: float | None = None                # absolute floor of the cosine schedule (sets min_lr_ratio = min_lr / lr)
    override_optimizer_config: dict = field(default_factory=dict)   # Megatron optimizer knobs (cpu offload fractions, ...): accepted;
                                                                    # ``optimizer_cpu_offload`` maps to fsdp_config.optimizer_offload

    def __post_init__(self):
        style = self.lr_decay_style or self.lr_scheduler_type
        if style:
            self.warmup_style = "cosine" if style == "cosine" else "constant"
        if self.grad_clip is not None:
            self.clip_grad = float(self.grad_clip)
        if self.lr_decay_steps and self.total_training_steps <= 0:
            self.total_training_steps = int(self.lr_decay_steps)
        if self.min_lr is not None and self.lr > 0:
            self.min_lr_ratio = float(self.min_lr) / float(self.lr)


@dataclass
class FSDPConfig(BaseConfig):
    param_offload: bool = False
    optimizer_offload: bool = False
    fsdp_size: int = -1
    reshard_after_forward: bool = True
    forward_prefetch: bool = True
    comm_backend: str = "p2p"           # "p2p" (NVLink peer-memory kernels) | "nccl" (baseline)
    model_dtype: str = "bf16"
    # keep a full bf16 replica on every GPU, re-gathered once per optimizer step and aliased by the colocated rollout
    # engine ("auto": when the replica is <= 1/4 of HBM; reshard_after_forward=False in FSDP terms)
    resident_params: str = "auto"       # auto | true | false
    # full = recompute every block in backward; none = save all activations; auto = save what fits in free HBM; or an int
    activation_checkpoint: str = "auto"
    activation_offload: bool = False     # park saved activations in pinned host memory (verl enable_activation_offload); slower than
                                         # recompute on B200, for parity and for models whose saved activations do not fit
    activation_reserve_gb: float = 16.0
    # accepted for config compatibility with torch-FSDP based stacks; the engine's precision policy is fixed (bf16 parameters and
    # gradients on the wire, fp32 master weights / moments / gradient accumulation) and parameters are always "original"
    mixed_precision: dict = field(default_factory=dict)
    use_orig_params: bool = True
    wrap_policy: dict = field(default_factory=dict)
    offload_policy: bool = False
    strategy: str = "fsdp"
    forward_only: bool = False
    ulysses_sequence_parallel_size: int | None = None
    entropy_from_logits_with_chunking: bool = False
    entropy_checkpointing: bool = False
    use_torch_compile: bool = False


@dataclass
class PolicyLossConfig(BaseConfig):
    loss_mode: str = "vanilla"
    clip_cov_ratio: float = 0.0002
    clip_cov_lb: float = 1.0
    clip_cov_ub: float = 5.0
    kl_cov_ratio: float = 0.0002
    ppo_kl_coef: float = 0.1
    sppo_eta: float = 1.0                # loss_mode=sppo: scale of the sequence reward the log-ratio is regressed onto
    dpo_beta: float = 0.1                # loss_mode=online_dpo (SPIN-style pairs, rollout.n = 2)
    dpo_label_smoothing: float = 0.0
    dpo_loss_type: str = "sigmoid"       # sigmoid | ipo
    flowrl_beta: float = 15.0            # loss_mode=flowrl: reward scale of the trajectory-balance target
    flowrl_logz_layers: int = 3          # depth of the log-partition head on the mean prompt hidden state
    flowrl_logz_dropout: float = 0.1
    flowrl_logz_lr: float = 1e-4


@dataclass
class CheckpointConfig(BaseConfig):
    save_contents: list = field(default_factory=lambda: ["model", "optimizer", "extra"])
    load_contents: list = field(default_factory=lambda: ["model", "optimizer", "extra"])
    async_save: bool = False


@dataclass
class ProfilerConfig(BaseConfig):
    """Per-role profiler switch (reference: verl/utils/profiler/config.py ProfilerConfig): which ranks of this role record, and
    tool options merged over ``global_profiler.global_tool_config`` — ``{"nsys": {"discrete": bool}, "torch": {"step_start": int,
    "step_end": int}, "torch_memory": {"trace_alloc_max_entries": int, "stack_depth": int}}``."""
    enable: bool = False
    all_ranks: bool = False
    ranks: list = field(default_factory=list)
    tool_config: dict = field(default_factory=dict)
    tool: str | None = None              # per-role tool (defaults to global_profiler.tool)
    save_path: str | None = None


@dataclass
class MegatronConfig(BaseConfig):
    """Model-parallel layout of a ``strategy=megatron`` role (reference keys: actor.megatron.{tensor,pipeline,virtual_pipeline}_model_
    parallel_size, verl/trainer/config/actor/megatron_actor.yaml).  Served by the stand-alone engines ``parallel/pipeline.py`` and
    ``parallel/tp_train.py`` through the single-controller trainer (``trainer/controller_ppo.py::PipelinePPOWorker``)."""
    tensor_model_parallel_size: int = 1
    pipeline_model_parallel_size: int = 1
    virtual_pipeline_model_parallel_size: int | None = None
    expert_model_parallel_size: int = 1            # MoE: experts split over this many workers (= all workers), dense part data-parallel
    # the rest of the reference's Megatron block is accepted so that its launch scripts parse; what has a counterpart is mapped
    # (offload flags -> the ZeRO engine's offload options), transformer-engine fusion switches have nothing to switch here
    context_parallel_size: int = 1                 # > 1 is rejected by validate_config (Ulysses SP covers long sequences)
    expert_tensor_parallel_size: int | None = None
    sequence_parallel: bool = True
    param_offload: bool = False
    grad_offload: bool = False
    optimizer_offload: bool = False
    use_distributed_optimizer: bool = True
    use_dist_checkpointing: bool = False
    dist_checkpointing_path: str | None = None
    use_mbridge: bool = False
    seed: int = 42
    override_transformer_config: dict = field(default_factory=dict)
    override_ddp_config: dict = field(default_factory=dict)
    override_mcore_model_config: dict = field(default_factory=dict)
    forward_only: bool = False


@dataclass
class ActorConfig(BaseConfig):
    strategy: str = "fsdp"
    megatron: MegatronConfig = field(default_factory=MegatronConfig)
    freeze_vision_tower: bool | None = None   # alias of model.freeze_vision_tower (verl spells it on the actor)
    # older spellings of policy_loss.* directly on the actor (entropy-mechanism recipe scripts)
    loss_mode: str | None = None
    clip_cov_ratio: float | None = None
    clip_cov_lb: float | None = None
    clip_cov_ub: float | None = None
    kl_cov_ratio: float | None = None
    ppo_kl_coef: float | None = None
    # logits-memory knobs of HF-model stacks: the fused linear-cross-entropy kernel never materialises [T, V] logits, so there is
    # nothing to chunk or checkpoint; torch.compile is not used
    entropy_checkpointing: bool = False
    entropy_from_logits_with_chunking: bool = False
    use_torch_compile: bool = False
    use_fused_kernels: bool = True            # always: the fused sm_100a kernels are the implementation
    use_remove_padding: bool = True           # always: batches are packed
    data_loader_seed: int | None = None
    load_weight: bool = True
    ppo_mini_batch_size: int = 256
    ppo_micro_batch_size: int | None = None
    ppo_micro_batch_size_per_gpu: int | None = None
    use_dynamic_bsz: bool = False
    ppo_max_token_len_per_gpu: int = 16384
    clip_ratio: float = 0.2
    clip_ratio_low: float = 0.2
    clip_ratio_high: float = 0.2
    clip_ratio_c: float = 3.0
    policy_loss: PolicyLossConfig = field(default_factory=PolicyLossConfig)
    loss_agg_mode: str = "token-mean"
    entropy_coeff: float = 0.0
    router_aux_loss_coef: float = 0.0    # MoE policies: weight of the router load-balancing loss (0 = off; HF default 0.001)
    use_kl_loss: bool = False
    kl_loss_coef: float = 0.001
    kl_loss_type: str = "low_var_kl"
    ppo_epochs: int = 1
    shuffle: bool = False
    ulysses_sequence_parallel_size: int = 1
    grad_clip: float = 1.0
    optim: OptimConfig = field(default_factory=OptimConfig)
    fsdp_config: FSDPConfig = field(default_factory=FSDPConfig)
    checkpoint: CheckpointConfig = field(default_factory=CheckpointConfig)
    profiler: ProfilerConfig = field(default_factory=ProfilerConfig)
    _mutable_fields = {"ppo_mini_batch_size", "ppo_micro_batch_size", "ppo_micro_batch_size_per_gpu"}

    def __post_init__(self):
        if self.loss_mode:
            self.policy_loss.loss_mode = self.loss_mode
        for k in ("clip_cov_ratio", "clip_cov_lb", "clip_cov_ub", "kl_cov_ratio", "ppo_kl_coef"):
            if getattr(self, k) is not None:
                setattr(self.policy_loss, k, float(getattr(self, k)))


@dataclass
class RolloutManagerConfig(BaseConfig):
    port: int = 5000
    endpoint: str | None = None
    config_path: str | None = None
    max_local_gen_s: float = 150.0


@dataclass
class MultiTurnConfig(BaseConfig):
    """Multi-turn / tool rollouts (reference: verl/workers/config/rollout.py MultiTurnConfig)."""
    enable: bool = False
    max_assistant_turns: int = 4
    max_user_turns: int = 4
    max_parallel_calls: int = 1
    max_tool_response_length: int = 256
    tool_response_truncate_side: str = "middle"
    tool_config_path: str | None = None
    interaction_config_path: str | None = None
    format: str = "hermes"
    use_inference_chat_template: bool = False
    tokenization_sanity_check_mode: str = "strict"   # strict | ignore_strippable | disable
    num_repeat_rollouts: int | None = None   # collabllm-style repeated simulated-user rollouts per turn (accepted; single rollout here)


@dataclass
class RolloutConfig(BaseConfig):
    name: str = "native"                # native | native-disaggregated | sglang-disaggregated (alias) | hf | naive
    mode: str = "sync"
    n: int = 1
    temperature: float = 1.0
    top_k: int = -1
    top_p: float = 1.0
    do_sample: bool = True
    prompt_length: int = 512
    response_length: int = 512
    dtype: str = "bfloat16"
    weight_dtype: str = "bf16"          # bf16 | fp8 (row-scaled e4m3 rollout weights)
    gpu_memory_utilization: float = 0.5
    free_cache_engine: bool = True
    tensor_model_parallel_size: int = 1
    data_parallel_size: int = 1
    pipeline_model_parallel_size: int = 1
    expert_parallel_size: int = 1
    max_num_batched_tokens: int = 16384
    max_num_seqs: int = 256
    max_model_len: int | None = None
    enable_chunked_prefill: bool = True
    enable_prefix_caching: bool = False     # cross-request KV page reuse (always on for multi-turn rollouts)
    kv_vmm: bool = False                    # KV pool in a CUDA-VMM range: release/resume keep addresses (graphs survive)
    load_format: str = "dummy"
    calculate_log_probs: bool = False
    log_prob_micro_batch_size_per_gpu: int | None = None
    log_prob_use_dynamic_bsz: bool = False
    log_prob_max_token_len_per_gpu: int = 16384
    min_stream_batch_size: int = 16
    cuda_graph: bool = True
    enforce_eager: bool | None = None       # vLLM / SGLang spelling: True = no CUDA graphs (sets cuda_graph = not enforce_eager)
    log_prob_micro_batch_size: int | None = None     # GLOBAL micro-batch (deprecated in verl): divided by the GPU count at start-up
    engine_kwargs: dict = field(default_factory=dict)      # back-end specific extras (vllm / sglang sections): accepted, unused by the native engine
    layered_summon: bool = False            # FSDP-to-engine weight gathering granularity: the ZeRO engine always gathers unit by unit
    multi_stage_wake_up: bool = False       # weights and KV cache already resume separately (release_memory / resume_memory tags)
    disable_log_stats: bool = True
    layer_name_map: dict = field(default_factory=dict)
    ignore_eos: bool = False                 # generate up to the response length regardless of EOS (benchmarking)
    cudagraph_capture_sizes: list | None = None   # batch sizes to capture decode graphs for (default: the engine's buckets)
    profiler: ProfilerConfig = field(default_factory=ProfilerConfig)
    skip_tokenizer_init: bool = True        # the engine works on token ids; text never enters it
    over_sample_rate: float = 0.0           # SGLang async over-sampling knob: the manager's continuation makes stragglers cheap instead
    update_weights_bucket_megabytes: int = 512   # bucket size of IPC weight hand-off in other stacks; weights are aliased / peer-pushed here
    agent: dict = field(default_factory=dict)    # {"agent_loop_config_path": ..., "num_workers": ...}
    trace: dict = field(default_factory=dict)    # {"backend": weave | mlflow | None, "token2text": bool} -> utils/rollout_trace.py
    skip_rollout: bool = False
    skip_dump_dir: str = "/tmp/rollout_dump"
    rollout_manager: RolloutManagerConfig = field(default_factory=RolloutManagerConfig)
    multi_turn: MultiTurnConfig = field(default_factory=MultiTurnConfig)
    val_kwargs: dict = field(default_factory=lambda: {"n": 1, "temperature": 0.0, "do_sample": False})

    def __post_init__(self):
        if self.pipeline_model_parallel_size > 1:
            raise NotImplementedError("rollout pipeline parallelism is not supported")
        if self.expert_parallel_size > 1 and self.expert_parallel_size != self.tensor_model_parallel_size * self.data_parallel_size:
            raise ValueError("expert_parallel_size must equal tensor_model_parallel_size * data_parallel_size")
        if self.name == "sglang-disaggregated":
            self.name = "native-disaggregated"
        if self.enforce_eager is not None:
            self.cuda_graph = not bool(self.enforce_eager)


@dataclass
class RefConfig(BaseConfig):
    strategy: str = "fsdp"
    log_prob_micro_batch_size: int | None = None      # GLOBAL micro-batch (deprecated spelling): divided by the GPU count at start-up
    ulysses_sequence_parallel_size: int | None = None  # the reference policy shares the actor's SP groups (must match if set)
    megatron: MegatronConfig = field(default_factory=MegatronConfig)
    profiler: ProfilerConfig = field(default_factory=ProfilerConfig)
    entropy_checkpointing: bool = False
    entropy_from_logits_with_chunking: bool = False
    use_torch_compile: bool = False
    load_weight: bool = True
    model: dict = field(default_factory=dict)   # reference policy from a different checkpoint: {"path": ...}
    log_prob_micro_batch_size_per_gpu: int | None = None
    log_prob_use_dynamic_bsz: bool = False
    log_prob_max_token_len_per_gpu: int = 16384
    fsdp_config: FSDPConfig = field(default_factory=FSDPConfig)


@dataclass
class ModelCfg(BaseConfig):
    path: str = "tiny-llama"
    use_remove_padding: bool = True
    enable_gradient_checkpointing: bool = True
    use_fused_kernels: bool = True
    lora_rank: int = 0
    lora_alpha: int = 16
    target_modules: str = "all-linear"
    freeze_vision_tower: bool = False     # vision-language models: train the text stack only (verl model.freeze_vision_tower)
    override_config: dict = field(default_factory=dict)
    tokenizer_path: str | None = None
    use_shm: bool = False                 # copy the checkpoint directory to /dev/shm once per node before loading (verl model.use_shm)
    trust_remote_code: bool = False       # no remote code is ever executed: models are this framework's own implementations
    external_lib: str | None = None       # python module imported before model construction (custom architectures in HF stacks)
    hf_config_path: str | None = None     # config directory when it differs from the weights directory
    custom_chat_template: str | None = None   # Jinja chat template that replaces the tokenizer's own
    use_liger: bool = False               # third-party fused kernels of other stacks: the own fused kernels are always on
    exclude_modules: str | None = None    # LoRA: regex of module names that get no adapter (e.g. ".*visual.*")
    lora_adapter_path: str | None = None  # LoRA: peft adapter directory to start from
    fused_kernel_options: dict = field(default_factory=dict)   # {"impl_backend": ...}: the fused path is the own sm_100a kernels
    input_tokenizer: str | None = None    # reward models whose tokenizer differs from the policy's (re-tokenise before scoring)
    fsdp_config: FSDPConfig = field(default_factory=FSDPConfig)   # critic / reward model: verl nests their fsdp_config under model
    # PRIME's implicit process reward model (algos/prime.py::ImplicitPRM): beta of the log-ratio reward, own optimizer, and when it
    # is updated relative to scoring the batch (none | before | after)
    beta_train: float = 0.05
    update: str = "none"
    optim: OptimConfig = field(default_factory=OptimConfig)
    ref_path: str | None = None
    enable_activation_offload: bool = False     # -> fsdp_config.activation_offload of the role (both verl spellings are accepted)
    enable_activation_offloading: bool = False


@dataclass
class ActorRolloutRefConfig(BaseConfig):
    hybrid_engine: bool = True
    nccl_timeout: int = 600
    model: ModelCfg = field(default_factory=ModelCfg)
    actor: ActorConfig = field(default_factory=ActorConfig)
    rollout: RolloutConfig = field(default_factory=RolloutConfig)
    ref: RefConfig = field(default_factory=RefConfig)


@dataclass
class CriticConfig(BaseConfig):
    enable: bool | None = None
    strategy: str = "fsdp"
    model: ModelCfg = field(default_factory=ModelCfg)
    optim: OptimConfig = field(default_factory=lambda: OptimConfig(lr=1e-5))
    ppo_mini_batch_size: int = 256
    ppo_micro_batch_size_per_gpu: int | None = None
    use_dynamic_bsz: bool = False
    ppo_max_token_len_per_gpu: int = 32768
    forward_max_token_len_per_gpu: int = 32768
    cliprange_value: float = 0.5
    loss_agg_mode: str = "token-mean"
    grad_clip: float = 1.0
    ppo_epochs: int = 1
    fsdp_config: FSDPConfig = field(default_factory=FSDPConfig)
    checkpoint: CheckpointConfig = field(default_factory=CheckpointConfig)
    profiler: ProfilerConfig = field(default_factory=ProfilerConfig)
    megatron: MegatronConfig = field(default_factory=MegatronConfig)
    forward_micro_batch_size: int | None = None
    forward_micro_batch_size_per_gpu: int | None = None
    nccl_timeout: int = 600
    load_weight: bool = True
    data_loader_seed: int | None = None
    ulysses_sequence_parallel_size: int = 1    # the critic trains on the actor's SP groups (must equal the actor's setting)
    ppo_micro_batch_size: int | None = None    # GLOBAL micro-batch (deprecated spelling)
    shuffle: bool = False
    rollout_n: int | None = None

    def __post_init__(self):
        self.model.fsdp_config = self.fsdp_config     # verl spells critic.model.fsdp_config.*: one object behind both paths


@dataclass
class OverlongBufferConfig(BaseConfig):
    """DAPO soft over-long punishment (reference: verl/workers/reward_manager/dapo.py, reward_model.overlong_buffer.*): responses
    longer than ``max_response_length - len`` lose up to ``penalty_factor`` linearly over the last ``len`` tokens."""
    enable: bool = False
    len: int = 0
    penalty_factor: float = 1.0
    log: bool = False


@dataclass
class RewardModelConfig(BaseConfig):
    enable: bool = False
    reward_manager: str = "naive"       # naive | batch | dapo | prime
    overlong_buffer: OverlongBufferConfig = field(default_factory=OverlongBufferConfig)
    use_reward_loop: bool = False       # async per-sample reward loop (data/reward_loop.py; managers naive | dapo)
    reward_loop_concurrency: int = 64
    genrm_endpoints: list = field(default_factory=list)   # "host:port" engine servers judging responses (generative reward model)
    genrm_max_new_tokens: int = 64
    launch_reward_fn_async: bool = False
    model: ModelCfg = field(default_factory=ModelCfg)
    micro_batch_size_per_gpu: int | None = None
    micro_batch_size: int | None = None              # GLOBAL micro-batch (deprecated spelling)
    use_dynamic_bsz: bool = False
    forward_max_token_len_per_gpu: int = 32768
    strategy: str = "fsdp"
    param_offload: bool = False
    max_length: int | None = None
    ulysses_sequence_parallel_size: int = 1
    nccl_timeout: int = 600
    load_weight: bool = True
    megatron: MegatronConfig = field(default_factory=MegatronConfig)
    profiler: ProfilerConfig = field(default_factory=ProfilerConfig)
    # reward_kwargs.{max_resp_len, overlong_buffer_cfg.{enable,len,penalty_factor,log}}: the older spelling of overlong_buffer
    reward_kwargs: dict = field(default_factory=dict)
    sandbox_fusion: dict = field(default_factory=dict)     # {"url": ..., "max_concurrent": ..., "memory_limit_mb": ...}
    # generative / separately placed reward models (fapo, genrm recipes): pool size and engine options of the judge; the judge
    # here is any engine server behind ``genrm_endpoints``
    enable_resource_pool: bool = False
    n_gpus_per_node: int = 0
    nnodes: int = 0
    rollout: dict = field(default_factory=dict)
    mini_batch_size: int | None = None      # PRIME: mini-batch of the implicit process reward model's own update

    def __post_init__(self):
        ob = (self.reward_kwargs or {}).get("overlong_buffer_cfg")
        if ob:
            self.overlong_buffer = OverlongBufferConfig(enable=bool(ob.get("enable", False)), len=int(ob.get("len", 0)),
                                                        penalty_factor=float(ob.get("penalty_factor", 1.0)), log=bool(ob.get("log", False)))


@dataclass
class CustomRewardFunction(BaseConfig):
    path: str | None = None
    name: str = "compute_score"


@dataclass
class KLCtrlConfig(BaseConfig):
    type: str = "fixed"
    kl_coef: float = 0.001
    horizon: int = 10000
    target_kl: float = 0.1


@dataclass
class RolloutCorrectionConfig(BaseConfig):
    rollout_is: str | None = None        # None | token | sequence | geometric
    rollout_is_threshold: float = 2.0
    rollout_rs: str | None = None        # None | token | sequence | geometric
    rollout_rs_threshold: float | None = None
    rollout_rs_threshold_lower: float | None = None
    rollout_token_veto_threshold: float | None = None
    bypass_mode: bool = False
    bypass_old_logprob_for_rollout: bool = False     # older spelling of bypass_mode
    use_pure_rollout_correction: bool = False

    def __post_init__(self):
        if self.bypass_old_logprob_for_rollout:
            self.bypass_mode = True


@dataclass
class FilterGroupsConfig(BaseConfig):
    """DAPO dynamic sampling (reference recipe: 3rdparty/verl/recipe/dapo/dapo_ray_trainer.py filter_groups): prompt groups whose
    samples all got the same ``metric`` carry no learning signal under group-normalised advantages and are taken out of the loss."""
    enable: bool = False
    metric: str = "seq_reward"            # seq_reward | seq_final_reward | acc | score (per-sample scalar looked up in the batch)
    max_num_gen_batches: int = 0          # recipes.dapo: give up after this many generation batches per step (0 = keep trying);
                                          # the streaming trainer masks groups instead and never re-generates


@dataclass
class AlgorithmConfig(BaseConfig):
    gamma: float = 1.0
    lam: float = 1.0
    adv_estimator: str = "gae"
    norm_adv_by_std_in_grpo: bool = True
    use_kl_in_reward: bool = False
    kl_penalty: str = "kl"
    kl_ctrl: KLCtrlConfig = field(default_factory=KLCtrlConfig)
    rollout_correction: RolloutCorrectionConfig = field(default_factory=RolloutCorrectionConfig)
    filter_groups: FilterGroupsConfig = field(default_factory=FilterGroupsConfig)
    # PF-PPO (policy filtration, arXiv:2409.06957; reference: verl/trainer/ppo/core_algos.py compute_pf_ppo_reweight_data): before the
    # update the batch is resampled with replacement, weighting samples by their score (trainer/core.py::pf_ppo_resample)
    # PRIME (recipes/prime.py): weights of the outcome reward and of the implicit process reward in the advantage
    reward_gt_coef: float = 5.0
    reward_dpo_coef: float = 5.0
    # flat spellings of rollout_correction.* used by older configs (mapped by finalize_config)
    rollout_is: bool | str | None = None
    rollout_is_threshold: float | None = None
    rollout_rs: bool | str | None = None
    rollout_rs_threshold: float | None = None
    rollout_rs_threshold_lower: float | None = None
    rollout_token_veto_threshold: float | None = None
    bypass_old_logprob_for_rollout: bool | None = None
    use_pure_rollout_correction: bool | None = None
    use_pf_ppo: bool = False
    pf_ppo: dict = field(default_factory=lambda: {"reweight_method": "pow", "weight_pow": 2.0})


@dataclass
class DataConfig(BaseConfig):
    train_files: Any = None
    val_files: Any = None
    prompt_key: str = "prompt"
    reward_fn_key: str = "data_source"
    max_prompt_length: int = 512
    max_response_length: int = 512
    train_batch_size: int = 1024
    gen_batch_size: int | None = None      # recipes.dapo: prompts per generation batch (default: train_batch_size)
    val_batch_size: int | None = None
    shuffle: bool = True
    seed: int = 1
    filter_overlong_prompts: bool = False
    truncation: str = "error"
    return_raw_chat: bool = False
    image_key: str = "images"             # vision-language datasets: column with the images of a row (verl data.image_key)
    video_key: str = "videos"             # column with the videos of a row: lists of frames, or {"frames": [...], "fps": f}
    image_min_pixels: int | None = None   # resize budget of the image preprocessor (None: model default)
    image_max_pixels: int | None = None
    dataloader_num_workers: int = 0
    # custom sampler / data generator classes (verl data.sampler.{class_path,class_name}, data.datagen.{path,name}); an empty path
    # resolves the name inside polyrl_b200.data.dynamic (e.g. SuccessRateCurriculumSampler, MockDataGenerator)
    sampler: dict = field(default_factory=dict)
    datagen: dict = field(default_factory=dict)
    apply_chat_template_kwargs: dict = field(default_factory=dict)   # extra arguments of tokenizer.apply_chat_template (e.g. reasoning_effort)
    custom_cls: dict = field(default_factory=dict)      # {"path": file.py, "name": DatasetClass}: dataset class replacing RLHFDataset
    trust_remote_code: bool = False
    use_shm: bool = False
    validation_shuffle: bool = False
    tokenizer: str | None = None            # tokenizer path when it differs from the model's
    train_max_samples: int = -1             # use only the first N rows of the training / validation files (-1: all)
    val_max_samples: int = -1
    tool_config_path: str | None = None
    return_raw_input_ids: bool = False
    return_full_prompt: bool = False
    return_multi_modal_inputs: bool = True
    image_patch_size: int = 14
    filter_overlong_prompts_workers: int = 1
    # PRIME-style accuracy filtering (recipe/prime): keep prompts whose group accuracy lies inside the bounds, over-sample to refill
    filter_accuracy: bool = False
    accuracy_lower_bound: float = 0.0
    accuracy_upper_bound: float = 1.0
    oversample_factor: float = 1.0
    synthetic: bool = False               # synthetic GSM8K-shaped prompts (no network on the target box)
    synthetic_num_samples: int = 4096
    synthetic_prompt_len_range: tuple = (64, 192)


@dataclass
class GlobalProfilerConfig(BaseConfig):
    tool: str | None = None              # nsys | torch | torch_memory
    steps: list | None = None
    profile_continuous_steps: bool = False
    save_path: str = "outputs/profile"
    global_tool_config: dict = field(default_factory=dict)


@dataclass
class TrainerConfig(BaseConfig):
    total_epochs: int = 1
    total_training_steps: int | None = None
    project_name: str = "polyrl_b200"
    experiment_name: str = "run"
    logger: list = field(default_factory=lambda: ["console"])
    nnodes: int = 1
    n_gpus_per_node: int = 8
    use_legacy_worker_impl: str = "auto"      # verl's switch between its two worker implementations: there is one here
    log_freq: int = 1
    del_local_ckpt_after_load: bool = False
    ray_wait_register_center_timeout: int = 300
    max_ckpt_to_keep: int | None = None        # one limit for every role (sets max_actor_ckpt_to_keep / max_critic_ckpt_to_keep)
    ref_update_freq: int = -1                  # SPIN / self-play: copy the actor into the reference policy every k steps (-1: never)
    tensorboard_dir: str | None = None         # also honoured through the TENSORBOARD_DIR environment variable
    rl_logging_board_dir: str | None = None
    num_global_batch: int = 1                  # TransferQueue sizing of the reference stack (no TransferQueue service here)
    num_data_storage_units: int = 1
    num_data_controllers: int = 1
    save_freq: int = -1
    test_freq: int = -1
    val_before_train: bool = False
    val_only: bool = False
    balance_batch: bool = True
    critic_warmup: int = 0
    resume_mode: str = "auto"            # disable | auto | resume_path
    resume_from_path: str | None = None
    default_local_dir: str = "checkpoints"
    default_hdfs_dir: str | None = None
    max_actor_ckpt_to_keep: int | None = None
    max_critic_ckpt_to_keep: int | None = None
    esi_redundant_time: float = 0.0
    rollout_data_dir: str | None = None
    validation_data_dir: str | None = None
    log_val_generations: int = 0
    device: str = "cuda"
    controller_backend: str = "spmd"     # spmd | local | mp | ray
    n_rollout_gpus: int = 0              # >0 => disaggregated: dedicated rollout GPUs
    # PolyRL's hybrid mode: in a disaggregated job every trainer GPU ALSO hosts a local engine (aliasing the trainer's weights)
    # that generates during the manager's time-boxed local window, then hands its HBM back to training while the remote
    # instances finish the partial generations (reference: local SGLang per trainer + /register_local_rollout_instances)
    local_rollout: bool = False
    # PolyRL's elastic deployment: no in-job rollout GPUs needed — every rank trains AND serves a hybrid local engine behind the
    # manager, and external engines (``python -m polyrl_b200.rollout.launch_server`` on other hosts) join or leave at any time; they get
    # the weights over TCP (weight_transfer/agent.py)
    elastic_rollout: bool = False
    # one-step-off-policy (reference recipe: 3rdparty/verl/recipe/one_step_off_policy): the rollout GPUs generate batch k+1
    # (with the weights of step k) while the trainers update on batch k; weights are pushed once batch k+1 is complete
    one_step_off: bool = False


@dataclass
class PPOConfig(BaseConfig):
    data: DataConfig = field(default_factory=DataConfig)
    actor_rollout_ref: ActorRolloutRefConfig = field(default_factory=ActorRolloutRefConfig)
    critic: CriticConfig = field(default_factory=CriticConfig)
    reward_model: RewardModelConfig = field(default_factory=RewardModelConfig)
    custom_reward_function: CustomRewardFunction = field(default_factory=CustomRewardFunction)
    algorithm: AlgorithmConfig = field(default_factory=AlgorithmConfig)
    trainer: TrainerConfig = field(default_factory=TrainerConfig)
    global_profiler: GlobalProfilerConfig = field(default_factory=GlobalProfilerConfig)
    # dedicated rollout resource pool of the one-step-off / fully-async recipes (``rollout.nnodes``, ``rollout.n_gpus_per_node``):
    # finalize_config() turns it into trainer.n_rollout_gpus of the disaggregated deployment
    rollout: dict = field(default_factory=dict)
    # fully-async recipe block (recipe/fully_async_policy): what maps onto this framework's streaming trainer is applied by
    # finalize_config — staleness_threshold >= 1 -> trainer.one_step_off (generation runs one policy version ahead),
    # use_rollout_log_probs -> rollout_correction.bypass_mode (engine log-probs serve as old log-probs); partial rollouts are the
    # manager's token-level continuation and always on; parameter sync happens every optimizer step
    async_training: dict = field(default_factory=dict)
    ray_kwargs: dict = field(default_factory=dict)        # Ray start-up options of the reference stack: accepted, unused
    transfer_queue: dict = field(default_factory=dict)


def finalize_config(cfg: "PPOConfig") -> "PPOConfig":
    """Resolve aliases that span config sections (run by load_config after every override has been applied)."""
    t, arr = cfg.trainer, cfg.actor_rollout_ref
    if t.max_ckpt_to_keep is not None:
        t.max_actor_ckpt_to_keep = t.max_actor_ckpt_to_keep or t.max_ckpt_to_keep
        t.max_critic_ckpt_to_keep = t.max_critic_ckpt_to_keep or t.max_ckpt_to_keep
    if arr.actor.freeze_vision_tower is not None:
        arr.model.freeze_vision_tower = bool(arr.actor.freeze_vision_tower)
    if arr.model.enable_activation_offload or arr.model.enable_activation_offloading:
        arr.actor.fsdp_config.activation_offload = True
    if cfg.critic.model.enable_activation_offload or cfg.critic.model.enable_activation_offloading:
        cfg.critic.fsdp_config.activation_offload = True
    if (arr.actor.optim.override_optimizer_config or {}).get("optimizer_cpu_offload"):
        arr.actor.fsdp_config.optimizer_offload = True
    if arr.actor.strategy == "megatron" and arr.actor.megatron.optimizer_offload:
        arr.actor.fsdp_config.optimizer_offload = True
    rc = cfg.algorithm.rollout_correction
    for k in ("rollout_is", "rollout_is_threshold", "rollout_rs", "rollout_rs_threshold", "rollout_rs_threshold_lower",
              "rollout_token_veto_threshold"):
        v = getattr(cfg.algorithm, k)
        if v is not None:
            setattr(rc, k, v)
    if cfg.algorithm.bypass_old_logprob_for_rollout:
        rc.bypass_mode = True
    at = cfg.async_training or {}
    if float(at.get("staleness_threshold", 0) or 0) >= 1:
        t.one_step_off = True
    if at.get("use_rollout_log_probs"):
        cfg.algorithm.rollout_correction.bypass_mode = True
    pool = cfg.rollout or {}
    if pool.get("n_gpus_per_node") and not t.n_rollout_gpus:
        t.n_rollout_gpus = int(pool.get("nnodes", 1)) * int(pool["n_gpus_per_node"])
    ob = (cfg.reward_model.reward_kwargs or {}).get("overlong_buffer_cfg")
    if ob:                                                       # older spelling of reward_model.overlong_buffer.*
        cfg.reward_model.overlong_buffer = OverlongBufferConfig(enable=bool(ob.get("enable", False)), len=int(ob.get("len", 0)),
                                                                penalty_factor=float(ob.get("penalty_factor", 1.0)), log=bool(ob.get("log", False)))
    # GLOBAL micro-batch sizes (the deprecated verl spelling) -> per GPU, like verl's normalisation at worker start-up
    n_train = max(1, t.n_gpus_per_node * t.nnodes - t.n_rollout_gpus)
    for obj, glob, per in ((arr.rollout, "log_prob_micro_batch_size", "log_prob_micro_batch_size_per_gpu"),
                           (arr.ref, "log_prob_micro_batch_size", "log_prob_micro_batch_size_per_gpu"),
                           (cfg.critic, "ppo_micro_batch_size", "ppo_micro_batch_size_per_gpu"),
                           (cfg.reward_model, "micro_batch_size", "micro_batch_size_per_gpu")):
        if getattr(obj, glob, None) and getattr(obj, per, None) is None:
            setattr(obj, per, max(1, int(getattr(obj, glob)) // n_train))
    sf = cfg.reward_model.sandbox_fusion or {}
    if sf.get("url") and not os.environ.get("SANDBOX_FUSION_URL"):
        os.environ["SANDBOX_FUSION_URL"] = str(sf["url"])         # the code scorer reads the service address from here
    if t.tensorboard_dir and not os.environ.get("TENSORBOARD_DIR"):
        os.environ["TENSORBOARD_DIR"] = str(t.tensorboard_dir)
    return cfg


# ---------------------------------------------------------------------------------------------
# loading / overrides
# ---------------------------------------------------------------------------------------------
def _coerce(value: Any, target: Any) -> Any:
    """Parse a CLI string into the type of the current value."""
    if not isinstance(value, str):
        return value
    parsed = yaml.safe_load(value)
    if isinstance(target, float) and isinstance(parsed, (int, str)):
        try:
            return float(parsed)
        except (TypeError, ValueError):
            return parsed
    if isinstance(target, tuple) and isinstance(parsed, list):
        return tuple(parsed)
    return parsed


_HYDRA_KEYS = ("hydra", "defaults", "_target_", "_recursive_", "_convert_", "_partial_")


def _resolve_interpolations(node: Any, root: Mapping, depth: int = 0) -> Any:
    """``${a.b.c}`` / ``${oc.select:a.b.c,default}`` / ``${oc.env:VAR,default}`` of Hydra / OmegaConf YAML files, resolved against the
    file's own tree (``root``).  References that the file does not define (they point into Hydra's composed defaults) resolve to
    None, i.e. the value keeps this framework's default."""
    import re

    if isinstance(node, Mapping):
        return {k: _resolve_interpolations(v, root, depth) for k, v in node.items()}
    if isinstance(node, list):
        return [_resolve_interpolations(v, root, depth) for v in node]
    if not isinstance(node, str) or "${" not in node or depth > 8:
        return node

    def lookup(path: str):
        cur: Any = root
        for part in path.strip().split("."):
            if not isinstance(cur, Mapping) or part not in cur:
                return None
            cur = cur[part]
        return cur

    def one(expr: str):
        expr = expr.strip()
        if expr.startswith("oc.select:"):
            key, _, default = expr[len("oc.select:"):].partition(",")
            v = lookup(key)
            return yaml.safe_load(default) if v is None and default.strip() else v
        if expr.startswith("oc.env:"):
            key, _, default = expr[len("oc.env:"):].partition(",")
            return os.environ.get(key.strip(), yaml.safe_load(default) if default.strip() else None)
        return lookup(expr)

    m = re.fullmatch(r"\$\{([^{}]+)\}", node.strip())
    if m:
        return _resolve_interpolations(one(m.group(1)), root, depth + 1)
    out = re.sub(r"\$\{([^{}]+)\}", lambda mm: str(one(mm.group(1))), node)
    return _resolve_interpolations(out, root, depth + 1)


def _apply_dict(obj: Any, d: Mapping, path: str = "", lenient: bool = False) -> None:
    """``lenient`` (YAML files written for the reference stack): Hydra bookkeeping keys are skipped, unresolved interpolations keep
    the default, and keys this framework does not have are reported once instead of aborting the load."""
    for k, v in d.items():
        if k in _HYDRA_KEYS:
            continue
        if isinstance(obj, dict):
            obj[k] = dict(v) if isinstance(v, Mapping) else v
            continue
        if not hasattr(obj, k):
            if lenient:
                _UNKNOWN_YAML_KEYS.append(path + k)
                continue
            raise KeyError(f"unknown config key {path + k!r}")
        cur = getattr(obj, k)
        if lenient and isinstance(v, str) and v.strip() == "???":
            continue                                    # OmegaConf's mandatory-value marker: keep the default
        if v is None and lenient and cur is not None and not isinstance(cur, (str, int, float, bool)):
            continue                                    # "null" for a structured field: keep the default tree
        if is_dataclass(cur) and isinstance(v, Mapping):
            _apply_dict(cur, v, path + k + ".", lenient)
            if hasattr(cur, "__post_init__"):
                cur.__post_init__()
        elif isinstance(cur, dict) and isinstance(v, Mapping):
            cur.update({kk: vv for kk, vv in v.items() if kk not in _HYDRA_KEYS})
        else:
            if v is None and lenient and isinstance(cur, (int, float, bool)) and not isinstance(cur, str):
                continue                                # unresolved ${...}: keep the default
            setattr(obj, k, _coerce(v, cur) if not isinstance(v, Mapping) else dict(v))


_UNKNOWN_YAML_KEYS: list[str] = []


def apply_overrides(cfg: Any, overrides: list[str]) -> Any:
    """``a.b.c=value`` (set), ``+a.b.c=value`` (set, key may be new inside a dict field), ``~a.b.c`` (reset)."""
    for ov in overrides:
        ov = ov.strip()
        if not ov:
            continue
        if ov.startswith("~"):
            key, value, reset = ov[1:], None, True
        else:
            if "=" not in ov:
                raise ValueError(f"override {ov!r} is not of the form key=value")
            key, value = ov.split("=", 1)
            reset = False
        key = key.lstrip("+")
        parts = key.split(".")
        obj = cfg
        for i, p in enumerate(parts[:-1]):
            nxt = obj.get(p) if isinstance(obj, dict) else getattr(obj, p, None)
            if nxt is None and not isinstance(obj, dict):
                raise KeyError(f"unknown config key {'.'.join(parts[:i + 1])!r}")
            if nxt is None:
                nxt = obj.setdefault(p, {})
            obj = nxt
        leaf = parts[-1]
        if isinstance(obj, dict):
            obj[leaf] = None if reset else yaml.safe_load(value)
            continue
        if not hasattr(obj, leaf):
            raise KeyError(f"unknown config key {key!r}")
        cur = getattr(obj, leaf)
        if reset:
            default = {f.name: f for f in fields(obj)}[leaf]
            setattr(obj, leaf, default.default if default.default is not dataclasses.MISSING else default.default_factory())
        else:
            setattr(obj, leaf, _coerce(value, cur))
        if hasattr(obj, "__post_init__"):
            obj.__post_init__()
    return cfg


def load_config(yaml_paths: list[str] | str | None = None, overrides: list[str] | None = None,
                base: PPOConfig | None = None) -> PPOConfig:
    cfg = copy.deepcopy(base) if base is not None else PPOConfig()
    if isinstance(yaml_paths, str):
        yaml_paths = [yaml_paths]
    for p in yaml_paths or []:
        with open(p) as f:
            d = yaml.safe_load(f) or {}
        hydra_style = any(k in d for k in ("hydra", "defaults")) or "${" in Path(p).read_text()
        if hydra_style:
            d = _resolve_interpolations(d, d)
        del _UNKNOWN_YAML_KEYS[:]
        _apply_dict(cfg, d, lenient=hydra_style)
        if _UNKNOWN_YAML_KEYS and int(os.environ.get("RANK", 0)) == 0:
            print(f"[config] {p}: ignored {len(_UNKNOWN_YAML_KEYS)} key(s) without a counterpart here: "
                  f"{', '.join(_UNKNOWN_YAML_KEYS[:12])}{' ...' if len(_UNKNOWN_YAML_KEYS) > 12 else ''}", flush=True)
    apply_overrides(cfg, overrides or [])
    return finalize_config(cfg)


def config_to_flat_dict(cfg: Any, prefix: str = "") -> dict[str, Any]:
    out = {}
    d = dataclasses.asdict(cfg) if is_dataclass(cfg) else cfg
    for k, v in d.items():
        if isinstance(v, dict):
            out.update(config_to_flat_dict(v, prefix + k + "."))
        else:
            out[prefix + k] = v
    return out


def need_critic(cfg: PPOConfig) -> bool:
    if cfg.critic.enable is not None:
        return bool(cfg.critic.enable)
    return cfg.algorithm.adv_estimator == "gae"


def need_reference_policy(cfg: PPOConfig) -> bool:
    a = cfg.actor_rollout_ref.actor
    return cfg.algorithm.use_kl_in_reward or a.use_kl_loss or a.policy_loss.loss_mode == "flowrl"      # FlowRL's target uses pi_ref


def validate_config(cfg: PPOConfig, n_gpus: int | None = None) -> None:
    """Cross-field validation (reference: verl/utils/config.py:74-200 and
    rlboost/verl_stream/trainer/main_stream.py:372-389)."""
    t, a, r = cfg.trainer, cfg.actor_rollout_ref.actor, cfg.actor_rollout_ref.rollout
    n_gpus = n_gpus or t.n_gpus_per_node * t.nnodes
    n_train = n_gpus - t.n_rollout_gpus
    if n_train <= 0:
        raise ValueError("no trainer GPUs left after reserving rollout GPUs")
    real_bsz = cfg.data.train_batch_size * r.n
    if real_bsz % n_train != 0:
        raise ValueError(f"train_batch_size*n ({real_bsz}) must be divisible by the number of trainer GPUs ({n_train})")
    if a.ppo_epochs != 1:
        raise ValueError("the streaming trainer requires actor.ppo_epochs == 1")
    if a.strategy not in ("fsdp", "fsdp2", "zero3", "megatron"):
        raise NotImplementedError(f"actor strategy {a.strategy!r}: fsdp | fsdp2 | zero3 (streaming + controller trainers) or megatron "
                                  "(pipeline-parallel actor through trainer/controller_ppo.py)")
    if not a.use_dynamic_bsz:
        if a.ppo_micro_batch_size_per_gpu is None and a.ppo_micro_batch_size is None:
            raise ValueError("set actor.ppo_micro_batch_size_per_gpu (or enable use_dynamic_bsz)")
        mb = a.ppo_micro_batch_size_per_gpu or max(1, a.ppo_micro_batch_size // n_train)
        if r.min_stream_batch_size % (mb * n_train) != 0:
            raise ValueError(f"min_stream_batch_size ({r.min_stream_batch_size}) must be divisible by "
                             f"micro_batch_size_per_gpu*n_gpus ({mb * n_train})")
    if r.min_stream_batch_size % r.n != 0:
        raise ValueError("min_stream_batch_size must be divisible by rollout.n (whole GRPO groups per streamed batch)")
    if need_critic(cfg) and not a.use_dynamic_bsz:
        cm = cfg.critic.ppo_micro_batch_size_per_gpu
        if cm is not None and a.ppo_micro_batch_size_per_gpu is not None and cm != a.ppo_micro_batch_size_per_gpu:
            raise ValueError("with GAE the critic and actor micro-batch sizes must match in streaming mode")
    if a.loss_agg_mode not in ("token-mean", "seq-mean-token-sum", "seq-mean-token-mean", "seq-mean-token-sum-norm"):
        raise ValueError(f"invalid loss_agg_mode {a.loss_agg_mode}")
    if cfg.algorithm.use_kl_in_reward and a.use_kl_loss:
        print("[config] NOTICE: both in-reward KL and KL loss are enabled")
    if r.tensor_model_parallel_size < 1:
        raise ValueError("tensor_model_parallel_size must be >= 1")
