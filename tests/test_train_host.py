"""CPU: host-side logic of roko_b200.train (datasets, loop, early stopping, checkpoints, gradient
averaging) with an in-memory .hdf5 stand-in and a small stock-torch model in place of the CUDA one."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from roko_b200 import train as T
from roko_b200.synth import structured_windows
from tests import fake_h5


class TinyModel(nn.Module):
    """(B,200,90) codes -> (B,90,5) logits: per-column code histogram through one linear layer."""

    def __init__(self):
        super().__init__()
        self.fc = nn.Linear(12, 5)

    def forward(self, x):
        hist = torch.nn.functional.one_hot(x.long(), 12).float().mean(dim=1)      # (B,90,12)
        return self.fc(hist)


def _register(path, n, seed, groups=2):
    x, y = structured_windows(n, seed=seed, return_truth=True)
    pos = np.zeros((n, 90, 2), np.int64)
    cut = n // groups
    parts = [(f"ctg_{k}", "ctg", pos[k * cut:(k + 1) * cut if k + 1 < groups else n],
              x[k * cut:(k + 1) * cut if k + 1 < groups else n], y[k * cut:(k + 1) * cut if k + 1 < groups else n])
             for k in range(groups)]
    fake_h5.register(path, {"ctg": "ACGT"}, parts)
    return x, y


def test_datasets_agree_and_skip_meta_groups():
    x, y = _register("mem://train_a", 10, seed=5)
    lazy = T.TrainDataset("mem://train_a", transform=T.TrainToTensor(), h5=fake_h5)
    mem = T.InMemoryTrainDataset("mem://train_a", transform=T.TrainToTensor(), h5=fake_h5)
    assert len(lazy) == len(mem) == 10
    for i in (0, 4, 5, 9):
        xa, ya = lazy[i]
        xb, yb = mem[i]
        assert torch.equal(xa, xb) and torch.equal(ya, yb)
        assert xa.dtype == torch.uint8 and tuple(xa.shape) == (200, 90) and tuple(ya.shape) == (90,)
        assert np.array_equal(xa.numpy(), x[i]) and np.array_equal(ya.numpy(), y[i])


def test_training_loop_learns_checkpoints_and_stops(tmp_path):
    _register("mem://train_b", 64, seed=6)
    _register("mem://val_b", 32, seed=7)
    logs = []
    torch.manual_seed(0)
    model = TinyModel()
    hist = T.train("mem://train_b", str(tmp_path), "mem://val_b", mem=True, batch_size=16, epochs=6, lr=5e-2,
                   model=model, device="cpu", h5=fake_h5, log=logs.append, seed=1)
    assert hist["epochs"] == 6 and len(hist["val_acc"]) == 6
    assert hist["train_loss"][-1] < hist["train_loss"][0]
    assert hist["val_acc"][-1] > 0.5
    files = sorted(os.listdir(tmp_path))
    assert len(files) == 1 and files[0].startswith("rnn_model_") and "_acc=" in files[0]      # n_saved = 1
    assert os.path.join(str(tmp_path), files[0]) == hist["checkpoint"]
    sd = torch.load(hist["checkpoint"])
    assert set(sd) == set(model.state_dict())
    assert any(str(m).startswith("Val epoch: 1,") for m in logs)


def test_early_stopping_after_patience_evaluations(tmp_path):
    _register("mem://train_c", 16, seed=8)
    _register("mem://val_c", 16, seed=9)
    hist = T.train("mem://train_c", str(tmp_path), "mem://val_c", mem=False, batch_size=8, epochs=50, lr=0.0,
                   patience=3, model=TinyModel(), device="cpu", h5=fake_h5, log=lambda *_: None, seed=2)
    assert hist["epochs"] == 4                 # first evaluation sets the best, three more without improvement
    assert len(os.listdir(tmp_path)) == 1


def test_early_stopping_and_checkpoint_rules(tmp_path):
    es = T.EarlyStopping(2)
    assert [es.step(s) for s in (0.5, 0.6, 0.6, 0.55)] == [False, False, False, True]
    es = T.EarlyStopping(2)
    assert [es.step(s) for s in (0.5, 0.4, 0.6, 0.5, 0.5)] == [False, False, False, False, True]
    ck = T.BestCheckpoint(str(tmp_path))
    m = TinyModel()
    a = ck.step(0.5, m)
    assert a and os.path.exists(a)
    assert ck.step(0.4, m) is None and os.path.exists(a)
    b = ck.step(0.7, m)
    assert b and os.path.exists(b) and not os.path.exists(a)
    assert os.path.basename(b) == "rnn_model_3_acc=0.7.pth"


def test_running_average_is_ignites(tmp_path):
    _register("mem://train_d", 24, seed=10)
    hist = T.train("mem://train_d", str(tmp_path), None, mem=True, batch_size=8, epochs=1, lr=0.0,
                   model=TinyModel(), device="cpu", h5=fake_h5, log=lambda *_: None, seed=3)
    assert hist["val_acc"] == [] and hist["checkpoint"] is None and len(hist["train_loss"]) == 1


def test_cli_signature_matches_reference():
    with pytest.raises(SystemExit):
        T.main(["--help"])
    assert (T.BATCH_SIZE, T.EPOCHS, T.LR, T.PATIENCE) == (128, 100, 1e-4, 7)


def test_slab_loader_matches_dataset_rows_and_shuffles():
    x, y = _register("mem://train_e", 21, seed=12)
    ds = T.InMemoryTrainDataset("mem://train_e", transform=T.TrainToTensor(), h5=fake_h5)
    seq = T.SlabLoader(ds, 8)
    assert len(seq) == 3
    got = list(seq)
    assert [b[0].shape[0] for b in got] == [8, 8, 5]                    # the ragged last batch is kept
    assert np.array_equal(torch.cat([b[0] for b in got]).numpy(), x) and np.array_equal(torch.cat([b[1] for b in got]).numpy(), y)
    assert got[0][0].dtype == torch.uint8 and got[0][1].dtype == torch.int64
    g = torch.Generator().manual_seed(3)
    sh = T.SlabLoader(ds, 8, shuffle=True, generator=g)
    a = torch.cat([b[1] for b in sh]).numpy()
    b = torch.cat([b[1] for b in sh]).numpy()                            # a new permutation every epoch
    assert not np.array_equal(a, y) and not np.array_equal(a, b)
    y64 = y.astype(np.int64)
    assert sorted(map(bytes, a)) == sorted(map(bytes, y64)) == sorted(map(bytes, b))   # each epoch is a permutation of the rows
    pairs = {bytes(xx): bytes(yy) for xx, yy in zip(x, y64)}
    for bx, by in T.SlabLoader(ds, 8, shuffle=True, generator=g):        # rows stay paired with their labels
        for xx, yy in zip(bx.numpy(), by.numpy()):
            assert pairs[bytes(xx)] == bytes(yy)
