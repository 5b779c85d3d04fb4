"""Synthetic GSM8K-shaped prompt data and a verifiable token-level reward (no network on the target box).

Prompt lengths follow the GSM8K prompt-length profile (~100-200 tokens with the chat template; the
reference preprocesses GSM8K with the template at
3rdparty/verl/examples/data_preprocess/gsm8k.py:57-70).  Tokens are uniform over the vocabulary
(excluding special ids).  The synthetic reward is a deterministic function of the response token ids so
that GRPO groups get non-degenerate advantages with random-init weights: it plays the role of the
rule-based GSM8K scorer (3rdparty/verl/verl/utils/reward_score/gsm8k.py:20-72) in benchmarks.
"""
from __future__ import annotations

import numpy as np
import torch
from torch.utils.data import Dataset


class SyntheticPromptDataset(Dataset):
    def __init__(self, vocab_size: int, num_samples: int = 4096, prompt_len_range=(64, 192), max_prompt_length: int = 256,
                 pad_token_id: int = 0, seed: int = 1, data_source: str = "synthetic/gsm8k_shaped"):
        self.vocab_size, self.n = vocab_size, num_samples
        self.lo, self.hi = prompt_len_range
        self.max_prompt_length = max(max_prompt_length, self.hi)
        self.pad = pad_token_id
        self.seed = seed
        self.data_source = data_source

    def __len__(self) -> int:
        return self.n

    def __getitem__(self, i: int) -> dict:
        rng = np.random.default_rng(self.seed * 1000003 + i)
        ln = int(rng.integers(self.lo, self.hi + 1))
        low = min(16, self.vocab_size - 1)
        toks = rng.integers(low, self.vocab_size, size=ln)
        p = self.max_prompt_length
        ids = np.full(p, self.pad, dtype=np.int64)
        ids[p - ln:] = toks                          # left padding, like RLHFDataset
        mask = np.zeros(p, dtype=np.int64)
        mask[p - ln:] = 1
        pos = np.clip(np.cumsum(mask) - 1, 0, None)
        answer = int(toks.sum() % 7)
        return {
            "input_ids": torch.from_numpy(ids), "attention_mask": torch.from_numpy(mask),
            "position_ids": torch.from_numpy(pos), "raw_prompt_ids": toks.tolist(), "data_source": self.data_source,
            "reward_model": {"style": "rule", "ground_truth": str(answer)}, "extra_info": {"index": i}, "index": i,
        }


def synthetic_token_reward(response_ids, ground_truth: str, vocab_size: int) -> float:
    """1.0 iff (sum of response token ids) mod 7 equals the prompt-derived answer class parity bucket.

    With random policies this fires for roughly 1/2 of the samples (we compare mod 2 of the mod-7 class), so
    every GRPO group almost surely mixes rewards 0 and 1."""
    if len(response_ids) == 0:
        return 0.0
    s = int(np.asarray(response_ids, dtype=np.int64).sum())
    return 1.0 if (s % 7) % 2 == int(ground_truth) % 2 else 0.0
