"""Step metrics (names kept dashboard-compatible with the reference):
``compute_data_metrics`` / ``compute_timing_metrics`` / ``compute_throughout_metrics`` / validation bootstrap
(3rdparty/verl/verl/trainer/ppo/metric_utils.py:80-490) and ``reduce_metrics``."""
from __future__ import annotations

from collections import defaultdict
from typing import Any, Callable

import numpy as np
import torch

from ..protocol import DataProto


def reduce_metrics(metrics: dict[str, list | float]) -> dict[str, float]:
    out = {}
    for k, v in metrics.items():
        if isinstance(v, (list, tuple)):
            if "max" in k.split("/")[-1]:
                out[k] = float(np.max(v))
            elif "min" in k.split("/")[-1]:
                out[k] = float(np.min(v))
            else:
                out[k] = float(np.mean(v))
        else:
            out[k] = float(v)
    return out


def append_to_dict(dst: dict[str, list], src: dict[str, Any]) -> None:
    for k, v in src.items():
        dst.setdefault(k, []).extend(v if isinstance(v, list) else [v])


def _response_info(batch: DataProto) -> dict:
    r = batch.batch["responses"].shape[-1]
    mask = batch.batch["attention_mask"]
    resp_mask = batch.batch["response_mask"] if "response_mask" in batch.batch else mask[:, -r:]
    return {"response_mask": resp_mask.float(), "prompt_length": mask[:, :-r].sum(-1).float(),
            "response_length": resp_mask.sum(-1).float()}


def compute_data_metrics(batch: DataProto, use_critic: bool = True) -> dict[str, float]:
    b = batch.batch
    info = _response_info(batch)
    m = info["response_mask"].bool()
    seq_score = b["token_level_scores"].sum(-1)
    seq_reward = b["token_level_rewards"].sum(-1)
    adv, ret = b["advantages"][m], b["returns"][m]
    max_resp = b["responses"].shape[-1]
    max_prompt = b["attention_mask"].shape[-1] - max_resp
    rl, pl = info["response_length"], info["prompt_length"]
    aborted = batch.non_tensor_batch.get("aborted")
    ok = torch.ones_like(rl, dtype=torch.bool) if aborted is None else ~torch.as_tensor(np.asarray(aborted, dtype=bool), device=rl.device)
    out = {
        "critic/score/mean": seq_score.mean().item(), "critic/score/max": seq_score.max().item(), "critic/score/min": seq_score.min().item(),
        "critic/rewards/mean": seq_reward.mean().item(), "critic/rewards/max": seq_reward.max().item(), "critic/rewards/min": seq_reward.min().item(),
        "critic/advantages/mean": adv.mean().item() if adv.numel() else 0.0, "critic/advantages/max": adv.max().item() if adv.numel() else 0.0,
        "critic/advantages/min": adv.min().item() if adv.numel() else 0.0,
        "critic/returns/mean": ret.mean().item() if ret.numel() else 0.0, "critic/returns/max": ret.max().item() if ret.numel() else 0.0,
        "critic/returns/min": ret.min().item() if ret.numel() else 0.0,
        "response_length/mean": rl.mean().item(), "response_length/max": rl.max().item(), "response_length/min": rl.min().item(),
        "response_length/clip_ratio": (rl == max_resp).float().mean().item(),
        "response_length_non_aborted/mean": rl[ok].mean().item() if ok.any() else 0.0,
        "response/aborted_ratio": 1.0 - ok.float().mean().item(),
        "prompt_length/mean": pl.mean().item(), "prompt_length/max": pl.max().item(), "prompt_length/min": pl.min().item(),
        "prompt_length/clip_ratio": (pl == max_prompt).float().mean().item(),
    }
    if use_critic and "values" in b:
        v = b["values"][m]
        var = torch.var(ret) if ret.numel() > 1 else torch.tensor(0.0)
        out.update({"critic/values/mean": v.mean().item(), "critic/values/max": v.max().item(), "critic/values/min": v.min().item(),
                    "critic/vf_explained_var": (1.0 - torch.var(ret - v) / (var + 1e-5)).item() if ret.numel() > 1 else 0.0})
    return out


def compute_timing_metrics(batch: DataProto, timing_raw: dict[str, float]) -> dict[str, float]:
    info = _response_info(batch)
    n_prompt = info["prompt_length"].sum().item()
    n_resp = info["response_length"].sum().item()
    tokens = {"gen": n_resp, **{k: n_prompt + n_resp for k in ("ref", "values", "adv", "update_critic", "update_actor", "old_log_prob")}}
    out = {f"timing_s/{k}": v for k, v in timing_raw.items()}
    out.update({f"timing_per_token_ms/{k}": timing_raw[k] * 1000 / tokens[k] for k in tokens if k in timing_raw and tokens[k] > 0})
    return out


def compute_throughout_metrics(batch: DataProto, timing_raw: dict[str, float], n_gpus: int) -> dict[str, float]:
    total = sum(batch.meta_info.get("global_token_num", [int(batch.batch["attention_mask"].sum().item())]))
    t = timing_raw["step"]
    return {"perf/total_num_tokens": total, "perf/time_per_step": t, "perf/throughput": total / (t * n_gpus),
            "perf/throughput_all_gpus": total / t}


def bootstrap_metric(data: list, subset_size: int, reduce_fns: list[Callable], n_bootstrap: int = 1000, seed: int = 42):
    rng = np.random.RandomState(seed)
    res = [[] for _ in reduce_fns]
    for _ in range(n_bootstrap):
        idx = rng.choice(len(data), size=subset_size, replace=True)
        sub = [data[i] for i in idx]
        for i, fn in enumerate(reduce_fns):
            res[i].append(fn(sub))
    return [(float(np.mean(r)), float(np.std(r))) for r in res]


def process_validation_metrics(data_sources: list[str], sample_uids: list[str], infos: dict[str, list], seed: int = 42) -> dict:
    """data_source -> var -> metric -> value with mean@N, best@N / worst@N (bootstrapped) and maj@N when
    predictions are available (reference: metric_utils.py:305-490)."""
    by = defaultdict(lambda: defaultdict(lambda: defaultdict(list)))
    for i, (ds, uid) in enumerate(zip(data_sources, sample_uids)):
        for var, vals in infos.items():
            by[ds][uid][var].append(vals[i])
    out = defaultdict(lambda: defaultdict(dict))
    for ds, uids in by.items():
        agg = defaultdict(lambda: defaultdict(list))
        for uid, vars_ in uids.items():
            for var, vals in vars_.items():
                if isinstance(vals[0], str):
                    continue
                n = len(vals)
                agg[var][f"mean@{n}"].append(float(np.mean(vals)))
                if n > 1:
                    (best, _), (worst, _) = bootstrap_metric(vals, n, [np.max, np.min], n_bootstrap=100, seed=seed)
                    agg[var][f"best@{n}/mean"].append(best)
                    agg[var][f"worst@{n}/mean"].append(worst)
                    if "pred" in vars_:
                        pairs = list(zip(vars_["pred"], vals))
                        counts = defaultdict(list)
                        for p, v in pairs:
                            counts[p].append(v)
                        maj = max(counts.items(), key=lambda kv: len(kv[1]))[1][0]
                        agg[var][f"maj@{n}/mean"].append(float(maj))
        for var, ms in agg.items():
            for mname, vals in ms.items():
                out[ds][var][mname] = float(np.mean(vals))
    return out
