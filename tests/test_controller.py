"""Single controller on CPU: dispatch / collect modes, auto padding, nd-mesh lazy dispatch, colocated (fused)
workers, non-blocking futures, multi-process backend with gloo collectives, aliveness check.
Mirrors the reference's Ray-on-CPU tier (3rdparty/verl/tests/single_controller/*)."""
import os
import signal
import subprocess
import sys
import textwrap

import numpy as np
import pytest
import torch

from polyrl_b200.controller import (ClassWithInitArgs, Dispatch, Execute, LocalWorkerGroup, MPWorkerGroup, Worker,
                                    create_colocated_worker_cls, make_nd_compute_dataproto_dispatch_fn, register)
from polyrl_b200.protocol import DataProto


class Toy(Worker):
    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = scale
        # 2-way model parallel inside a dp group: ranks (0,1) -> dp 0, (2,3) -> dp 1; collect from even ranks
        self._register_dispatch_collect_info("actor", dp_rank=self.rank // 2, is_collect=self.rank % 2 == 0,
                                             dp_size=max(1, self.world_size // 2))

    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def whoami(self):
        return (self.rank, self.world_size)

    @register(dispatch_mode=Dispatch.RANK_ZERO, execute_mode=Execute.RANK_ZERO)
    def only_zero(self, x):
        return x + self.rank

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def double(self, data: DataProto):
        return DataProto.from_dict({"y": data.batch["x"] * 2 * self.scale, "r": torch.full((len(data),), self.rank)})

    @register(dispatch_mode=make_nd_compute_dataproto_dispatch_fn("actor"))
    def mesh_fn(self, data: DataProto):
        return DataProto.from_dict({"y": data.batch["x"] + 100 * self.rank})

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO, blocking=False)
    def slow_double(self, data: DataProto):
        import time
        time.sleep(0.05)
        return DataProto.from_dict({"y": data.batch["x"] * 2})


def test_dispatch_modes_local():
    wg = LocalWorkerGroup(4, ClassWithInitArgs(Toy, scale=1.0))
    assert wg.whoami() == [(i, 4) for i in range(4)]
    assert wg.only_zero(5) == 5
    data = DataProto.from_dict({"x": torch.arange(8.0)})
    out = wg.double(data)
    assert torch.equal(out.batch["y"], torch.arange(8.0) * 2)
    assert out.batch["r"].tolist() == [0, 0, 1, 1, 2, 2, 3, 3]
    # auto padding: 6 samples over 4 workers
    out = wg.double(DataProto.from_dict({"x": torch.arange(6.0)}))
    assert len(out) == 6 and torch.equal(out.batch["y"], torch.arange(6.0) * 2)
    # nd-mesh: dp_size 2, data replicated inside each pair, collected from ranks 0 and 2
    out = wg.mesh_fn(DataProto.from_dict({"x": torch.arange(4.0)}))
    assert out.batch["y"].tolist() == [0.0, 1.0, 202.0, 203.0]
    fut = wg.slow_double(data)
    assert torch.equal(fut.get().batch["y"], torch.arange(8.0) * 2)
    wg.shutdown()


class Critic(Worker):
    @register(dispatch_mode=Dispatch.ONE_TO_ALL)
    def name(self):
        return f"critic{self.rank}"


def test_colocated_workers():
    fused = create_colocated_worker_cls({"actor": ClassWithInitArgs(Toy, scale=0.5), "critic": ClassWithInitArgs(Critic)})
    wg = LocalWorkerGroup(2, fused)
    views = wg.spawn({"actor", "critic"})
    assert views["critic"].name() == ["critic0", "critic1"]
    out = views["actor"].double(DataProto.from_dict({"x": torch.ones(4)}))
    assert torch.equal(out.batch["y"], torch.ones(4))
    wg.shutdown()


class DistWorker(Worker):
    def __init__(self):
        super().__init__()
        import torch.distributed as dist
        dist.init_process_group("gloo", rank=self.rank, world_size=self.world_size)

    @register(dispatch_mode=Dispatch.DP_COMPUTE_PROTO)
    def allreduce_sum(self, data: DataProto):
        import torch.distributed as dist
        t = data.batch["x"].sum().reshape(1).clone()
        dist.all_reduce(t)
        return DataProto.from_dict({"s": t.repeat(len(data))}, non_tensors={"pid": np.array([os.getpid()] * len(data))})


def test_mp_backend_gloo():
    wg = MPWorkerGroup(2, ClassWithInitArgs(DistWorker))
    out = wg.allreduce_sum(DataProto.from_dict({"x": torch.arange(6.0)}))
    assert out.batch["s"].tolist() == [15.0] * 6
    assert len(set(out.non_tensor_batch["pid"].tolist())) == 2 and os.getpid() not in out.non_tensor_batch["pid"]
    wg.shutdown()


def test_worker_death_aborts_driver(tmp_path):
    script = tmp_path / "die.py"
    script.write_text(textwrap.dedent(f"""
        import sys, time
        sys.path.insert(0, {str(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))!r})
        from polyrl_b200.controller import ClassWithInitArgs, Dispatch, MPWorkerGroup, Worker, register
        class Dying(Worker):
            @register(dispatch_mode=Dispatch.ONE_TO_ALL, blocking=False)
            def die(self):
                import os; os._exit(1)
        if __name__ == "__main__":
            wg = MPWorkerGroup(2, ClassWithInitArgs(Dying))
            wg.start_worker_aliveness_check(0.1)
            wg.die()
            time.sleep(10)
            print("driver survived")
    """))
    r = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=120)
    assert r.returncode == -signal.SIGABRT, (r.returncode, r.stdout, r.stderr[-500:])
    assert "driver survived" not in r.stdout
