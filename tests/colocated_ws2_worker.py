"""The scaling-bench topology on CPU/gloo (2 ranks): colocated GRPO steps with per-rank rollout replicas, ZeRO-sharded fp32 state,
resident replica and saved activations (bench.py --gpus 2 runs the same ColocatedPPO.step on GPUs)."""
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dist.init_process_group("gloo")
from polyrl_b200.config.schema import load_config
from polyrl_b200.data.synthetic import SyntheticPromptDataset
from polyrl_b200.protocol import DataProto, collate_fn
from polyrl_b200.trainer.colocated import ColocatedPPO
rank = dist.get_rank()
cfg = load_config(overrides=["trainer.device=cpu", "actor_rollout_ref.model.path=tiny-llama", "algorithm.adv_estimator=grpo", "actor_rollout_ref.rollout.n=4",
    "data.max_prompt_length=32", "data.max_response_length=12", "data.train_batch_size=8", "actor_rollout_ref.actor.ppo_mini_batch_size=8",
    "actor_rollout_ref.actor.use_dynamic_bsz=True", "actor_rollout_ref.actor.ppo_max_token_len_per_gpu=256", "actor_rollout_ref.rollout.log_prob_use_dynamic_bsz=True",
    "actor_rollout_ref.rollout.log_prob_max_token_len_per_gpu=256", "actor_rollout_ref.rollout.max_num_seqs=16", "actor_rollout_ref.rollout.cuda_graph=False",
    "trainer.n_gpus_per_node=2"])
ppo = ColocatedPPO(cfg, device="cpu", group=None)
ds = SyntheticPromptDataset(ppo.worker.model_cfg.vocab_size, 64, (8, 24), 32)
for step in range(2):
    b = DataProto.from_single_dict(collate_fn([ds[(step * 2 + rank) * 4 + i] for i in range(4)]))
    m = ppo.step(b)
    if rank == 0: print("COLO2 step", step, "resident", ppo.worker.actor_engine.resident, "saved", ppo.worker.actor_engine.saved_layers_last, "gn", round(m["actor/grad_norm"], 4), "n", m["num_samples"], flush=True)
dist.barrier(); dist.destroy_process_group()
