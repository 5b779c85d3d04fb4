"""MoE decode on the GPU: engine log-probs of its own samples vs a full forward (run by scripts/gpu_round2_checks.sh, also with
POLYRL_MOE_STATIC_DECODE=1 = CUDA-graph decode over the static-shape expert path)."""
import torch
from polyrl_b200.models import transformer as tf
from polyrl_b200.models.config import get_model_config
from polyrl_b200.rollout.engine import RolloutEngine, SamplingParams
cfg = get_model_config("tiny-qwen3-moe-128")
w = tf.TransformerWeights(cfg, device="cuda", dtype=torch.bfloat16, seed=5)
eng = RolloutEngine(cfg, "cuda", max_num_seqs=8, max_model_len=128, kv_cache_tokens=8 * 128)
eng.set_weights(w.embed, w.layers, w.head)
out = eng.generate([[5, 6, 7, 8], [9, 10, 11]], SamplingParams(max_new_tokens=16, temperature=1.0, ignore_eos=True, seed=3))
for grp in out:
    r = grp[0]
    ids = torch.tensor(r.prompt_ids + r.output_ids, device="cuda")
    inp = tf.PackedInput(ids, torch.arange(ids.numel(), device="cuda"), torch.tensor([0, ids.numel()], dtype=torch.int32, device="cuda"), ids.numel())
    with torch.no_grad():
        lp = torch.log_softmax(tf.logits_forward(cfg, w.head, w.embed, tf.forward_hidden(cfg, w, inp)).float(), -1)
    k = len(r.prompt_ids) - 1
    got = torch.tensor(r.logprobs, device="cuda")
    ref = lp[k:k + len(r.output_ids)].gather(-1, ids[k + 1:].unsqueeze(-1)).squeeze(-1)
    assert (got - ref).abs().max() < 0.2, (got - ref).abs().max()
print("moe decode ok")
