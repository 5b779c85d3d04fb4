"""DataProto protocol (mirrors 3rdparty/verl/tests/test_protocol_on_cpu.py in spirit)."""
import pickle

import numpy as np
import pytest
import torch

from polyrl_b200.protocol import (DataProto, DataProtoFuture, TensorBatch, collate_fn, pad_dataproto_to_divisor,
                                  unpad_dataproto)


def _mk(n=6):
    return DataProto.from_dict({"a": torch.arange(n * 2).view(n, 2), "b": torch.arange(n).float()},
                               {"s": [f"x{i}" for i in range(n)]}, {"tag": 1})


def test_basic_and_consistency():
    d = _mk()
    assert len(d) == 6 and set(d.keys()) == {"a", "b", "s"}
    with pytest.raises(ValueError):
        DataProto.from_dict({"a": torch.zeros(3)}, {"s": ["x"] * 4})
    with pytest.raises(ValueError):
        TensorBatch({"a": torch.zeros(3), "b": torch.zeros(4)})
    assert d[2].batch["b"].tolist() == [2.0] and d[1:3].non_tensor_batch["s"].tolist() == ["x1", "x2"]


def test_chunk_split_concat_roundtrip():
    d = _mk()
    parts = d.chunk(3)
    assert [len(p) for p in parts] == [2, 2, 2]
    back = DataProto.concat(parts)
    assert torch.equal(back.batch["a"], d.batch["a"]) and back.non_tensor_batch["s"].tolist() == d.non_tensor_batch["s"].tolist()
    with pytest.raises(ValueError):
        d.chunk(4)
    assert [len(p) for p in d.split(4)] == [4, 2]


def test_union_select_pop_rename():
    d = _mk()
    e = DataProto.from_dict({"c": torch.ones(6)}, {"t": list(range(6))}, {"tag": 1, "k": 2})
    d.union(e)
    assert "c" in d.batch and d.meta_info["k"] == 2
    with pytest.raises(ValueError):
        d.union(DataProto.from_dict({"c": torch.zeros(6)}))
    s = d.select(batch_keys=["a"], non_tensor_batch_keys=["s"])
    assert list(s.batch.keys()) == ["a"] and list(s.non_tensor_batch) == ["s"]
    p = d.pop(batch_keys=["c"], non_tensor_batch_keys=["t"])
    assert "c" not in d.batch and "c" in p.batch and "t" not in d.non_tensor_batch
    d.rename("a", "aa")
    assert "aa" in d.batch and "a" not in d.batch


def test_repeat_reorder_index():
    d = _mk(3)
    r = d.repeat(2, interleave=True)
    assert r.batch["b"].tolist() == [0, 0, 1, 1, 2, 2] and r.non_tensor_batch["s"].tolist() == ["x0", "x0", "x1", "x1", "x2", "x2"]
    r2 = d.repeat(2, interleave=False)
    assert r2.batch["b"].tolist() == [0, 1, 2, 0, 1, 2]
    d.reorder(torch.tensor([2, 0, 1]))
    assert d.batch["b"].tolist() == [2, 0, 1] and d.non_tensor_batch["s"].tolist() == ["x2", "x0", "x1"]
    sel = d.select_idxs(torch.tensor([True, False, True]))
    assert sel.batch["b"].tolist() == [2, 1]
    assert d[[0, 2]].batch["b"].tolist() == [2, 1]


def test_padding_and_pickle_and_future():
    d = _mk(5)
    p, pad = pad_dataproto_to_divisor(d, 4)
    assert len(p) == 8 and pad == 3
    assert torch.equal(unpad_dataproto(p, pad).batch["a"], d.batch["a"])
    q = pickle.loads(pickle.dumps(d))
    assert torch.equal(q.batch["a"], d.batch["a"]) and q.non_tensor_batch["s"].tolist() == d.non_tensor_batch["s"].tolist() and q.meta_info == d.meta_info

    class F:
        def __init__(self, v): self.v = v
        def result(self): return self.v
    fut = DataProtoFuture.concat([F(x) for x in _mk(4).chunk(2)])
    assert len(fut.get()) == 4
    halves = fut.chunk(2)
    assert halves[1].get().batch["b"].tolist() == [2.0, 3.0]


def test_collate_and_iterator():
    rows = [{"x": torch.tensor([i, i]), "name": f"n{i}", "meta": {"i": i}} for i in range(4)]
    b = DataProto.from_single_dict(collate_fn(rows))
    assert b.batch["x"].shape == (4, 2) and b.non_tensor_batch["meta"][2] == {"i": 2}
    seen = [len(m) for m in b.make_iterator(2, epochs=2, seed=0)]
    assert seen == [2, 2, 2, 2]


def test_rollout_skip_replays_dumped_generations(tmp_path):
    import torch

    from polyrl_b200.protocol import DataProto
    from polyrl_b200.utils.rollout_skip import RolloutSkip

    calls = []

    class W:
        def generate_sequences(self, prompts):
            calls.append(1)
            return DataProto.from_dict({"input_ids": prompts.batch["input_ids"], "responses": torch.full((2, 3), len(calls))})

    p = DataProto.from_dict({"input_ids": torch.arange(8).view(2, 4)})
    w = W()
    RolloutSkip(str(tmp_path), n=2, train_batch_size=2).wrap_generate_sequences(w)
    first = w.generate_sequences(p)
    w2 = W()
    skip2 = RolloutSkip(str(tmp_path), n=2, train_batch_size=2)
    skip2.wrap_generate_sequences(w2)
    again = w2.generate_sequences(p)
    assert len(calls) == 1 and skip2.hits == 1 and torch.equal(first.batch["responses"], again.batch["responses"])
    w2.generate_sequences(p)                      # step 1 was never dumped: falls through to the engine
    assert len(calls) == 2
