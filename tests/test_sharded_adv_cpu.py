"""Group-level advantages with KL-in-reward when the streamed batch is sharded across DP ranks:
the shard-local rewards are gathered, the estimator runs on the whole batch, and the adaptive KL controller sees the
global mean — 2 gloo ranks must reproduce the single-process result row for row."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from polyrl_b200.algos import core_algos
from polyrl_b200.protocol import DataProto, pad_dataproto_to_divisor
from polyrl_b200.trainer import core


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _full_batch(n_groups=5, n=3, R=6, seed=0):
    g = torch.Generator().manual_seed(seed)
    rows = n_groups * n
    mask = (torch.rand(rows, R, generator=g) > 0.25).float()
    mask[:, 0] = 1
    scores = torch.zeros(rows, R)
    last = mask.cumsum(-1).argmax(-1)
    scores[torch.arange(rows), last] = torch.randn(rows, generator=g)
    old = -torch.rand(rows, R, generator=g)
    refp = old + 0.3 * torch.randn(rows, R, generator=g)
    uid = np.array([f"g{i // n}" for i in range(rows)], dtype=object)
    return DataProto.from_dict({"token_level_scores": scores, "response_mask": mask, "old_log_probs": old, "ref_log_prob": refp},
                               non_tensors={"uid": uid})


def _oracle(estimator):
    full = _full_batch()
    ctrl = core_algos.AdaptiveKLController(0.2, 0.1, 100)
    full, _ = core.apply_kl_penalty(full, ctrl, "kl")
    full = core.compute_advantage(full, estimator, 1.0, 1.0, 3, True, None)
    return full, ctrl.value


def _worker(rank, world, port, estimator, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        full = _full_batch()
        padded, pad = pad_dataproto_to_divisor(full, world)
        is_pad = np.zeros(len(padded), dtype=bool)
        if pad:
            padded.batch["response_mask"][-pad:] = 0
            is_pad[-pad:] = True
        order = list(reversed(range(len(padded))))           # any permutation that splits groups across ranks
        order = order[1::2] + order[0::2]
        padded.reorder(torch.tensor(order))
        is_pad = is_pad[np.asarray(order)]
        ctx = {"dp": world, "dp_rank": rank, "sp": 1, "is_pad": is_pad, "uid": padded.non_tensor_batch["uid"].copy()}
        local = padded.chunk(world)[rank]
        n = len(local)
        real = torch.from_numpy(~is_pad[rank * n:(rank + 1) * n])
        ctrl = core_algos.AdaptiveKLController(0.2, 0.1, 100)
        local, _ = core.apply_kl_penalty(local, ctrl, "kl", group=None, real_rows=real, distributed=True)
        local = core.compute_group_advantage_sharded(local, ctx, estimator, 1.0, 1.0, 3, True, None, None)
        rows = order[rank * n:(rank + 1) * n]
        out[rank] = (rows, local.batch["advantages"].clone(), ctrl.value, is_pad[rank * n:(rank + 1) * n].copy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("estimator", ["grpo", "rloo"])
def test_group_advantage_with_kl_in_reward_matches_unsharded(estimator):
    want, want_kl = _oracle(estimator)
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(2, _free_port(), estimator, out), nprocs=2, join=True)
        got = dict(out)
    assert abs(got[0][2] - got[1][2]) < 1e-12, "KL controllers diverged across ranks"
    assert abs(got[0][2] - want_kl) < 1e-6
    seen = 0
    for r in (0, 1):
        rows, adv, _, padflag = got[r]
        for j, src in enumerate(rows):
            if padflag[j]:
                continue
            seen += 1
            assert torch.allclose(adv[j], want.batch["advantages"][src], atol=1e-5), (r, j, src)
    assert seen == len(want)
