"""CPU: training-side plumbing that follows the hot path -- board symmetries against the reference's outputs,
replay store round trip, ReplayDataset position averaging, one tiny NeuralNetWrapper.train step with a
reference-format checkpoint."""
import os

import numpy as np
import pandas as pd
import pytest
import torch

from golden_io import load, unhex


def test_symmetries_match_reference():
    from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SymmetriesGenerator
    gen = SymmetriesGenerator()
    for rec in load("symmetries"):
        r = rec["L"] + 1
        b = torch.from_numpy(unhex(rec["boards"], np.float32).reshape(5, 3, r, r))
        p = torch.from_numpy(unhex(rec["policies"], np.float32).reshape(5, -1))
        for i, ref in enumerate(rec["out"]):
            ob, op = gen(b.clone(), p.clone(), index=i)
            assert np.array_equal(ob.numpy().ravel(), unhex(ref["boards"], np.float32)), (rec["L"], i)
            assert np.array_equal(op.numpy().ravel(), unhex(ref["policies"], np.float32)), (rec["L"], i)
    # non-square boards only have the four flips
    ob, op = gen(torch.zeros(2, 3, 3, 4), torch.rand(2, 24), index=6)
    assert ob.shape == (2, 3, 3, 4) and op.shape == (2, 24)


def _frame(n_games=6, rows=9, seed=0):
    rng = np.random.RandomState(seed)
    recs = []
    for g in range(n_games):
        for m in range(rows):
            x = rng.randint(0, 2, 48)
            pi = rng.rand(32); pi /= pi.sum()
            recs.append(dict(generation=np.int16(1), game_idx=np.int16(g), move_idx=np.int16(m), move=np.int16(m), player=np.int8(m % 2),
                             **{f"x_{i}": np.int16(x[i]) for i in range(48)}, **{f"pi_{i}": pi[i] for i in range(32)},
                             z=float(rng.choice([-1, 1])), max_deepness=np.int16(3), tree_size=np.int32(5), terminal_count=np.int32(0),
                             q_value=np.float32(0.1), training=np.int8(0)))
    return pd.DataFrame(recs).set_index(["generation", "game_idx", "move_idx"])


def test_replay_store_parquet_roundtrip(tmp_path):
    from dotsboxesaz_b200.utils.utils import ReplayDataset, ReplayStore
    store = ReplayStore(str(tmp_path / "sp_data"))
    assert not store.has("fresh")
    a, b = _frame(seed=1), _frame(seed=2)
    store.append("fresh", a)
    store.append("fresh", b)
    assert store.has("fresh")
    back = store.read("fresh")
    assert len(back) == len(a) + len(b) and list(back.columns) == list(a.columns)
    assert np.allclose(back["z"].to_numpy()[:len(a)], a["z"].to_numpy())
    store.delete("fresh")
    assert not store.has("fresh")
    # position averaging merges identical feature rows
    d = pd.concat([a, a]).assign(training=np.int8(1))
    ds = ReplayDataset(d, True, (3, 4, 4), pos_average=True)
    assert len(ds) <= len(a) and ds[0][0].shape == (3, 4, 4) and ds[0][1].shape == (32,)


def test_one_training_generation_and_checkpoint_format(tmp_path):
    from dotsboxesaz_b200.dots_boxes.dots_boxes_nn import SimpleNN, SymmetriesGenerator
    from dotsboxesaz_b200.nn import NeuralNetWrapper
    from dotsboxesaz_b200.utils.utils import DotDict, ReplayDataset
    params = DotDict({"nn": {"pytorch_device": "cpu", "chkpts_filename": str(tmp_path / "model_gen{}.pt"),
                             "train_params": {"symmetries": SymmetriesGenerator(), "nb_epochs": 1, "train_batch_size": 16,
                                              "val_batch_size": 16, "lr": 1e-2,
                                              "optimizer_params": {"momentum": 0.9, "weight_decay": 1e-4}}}})
    torch.manual_seed(0)
    model = SimpleNN(board=(3, 3))
    w0 = model.fc0.weight.detach().clone()
    df = _frame(n_games=8).assign(training=np.int8(1))
    ds = ReplayDataset(df, True, (3, 4, 4))
    wrapper = NeuralNetWrapper(model, params)
    # generation 1 resumes from generation 0's checkpoint, as the reference does (nn.py:194-196)
    from dotsboxesaz_b200.nn import save_checkpoint
    save_checkpoint(params.nn.chkpts_filename.format(0), model, torch.optim.SGD(model.parameters(), lr=1e-2, momentum=0.9), 0)
    last = wrapper.train(ds, None, None, 1)
    assert last == len(ds) // 16
    assert not torch.equal(w0, model.fc0.weight.detach())
    ck = torch.load(params.nn.chkpts_filename.format(1), map_location="cpu")
    assert set(ck) == {"last_batch_idx", "model_dict", "optimizer_dict"} and ck["last_batch_idx"] == last
