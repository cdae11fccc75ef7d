"""CPU: the oracle (numpy restatement + stock-torch port) against the reference-generated golden
vectors.  The fixtures were produced by EXECUTING /root/reference/roko/rnn_model.py
(oracle/make_golden.py); nothing here reads /root/reference."""
import numpy as np
import pytest
import torch

from oracle import roko_oracle as O
from oracle.torch_port import TorchCpuPort
from roko_b200.synth import structured_windows, uniform_windows


def test_state_dict_contract(seed1_state):
    assert list(seed1_state.keys()) == O.STATE_KEYS
    assert len(seed1_state) == 31
    for k, v in seed1_state.items():
        assert tuple(v.shape) == O.STATE_SHAPES[k] and v.dtype == torch.float32
    assert sum(v.numel() for v in seed1_state.values()) == 1099731


def test_oracle_f32_matches_reference(seed1_weights, golden):
    taps = {}
    logits = O.forward(golden["x"], seed1_weights, np.float32, taps)
    assert np.abs(logits - golden["logits"]).max() <= 2e-6
    assert np.array_equal(O.labels_from_logits(logits), golden["labels"])
    for k in ("front", "gru_l0", "gru_l1", "gru_l2"):
        assert np.abs(taps[k][:2] - golden["tap_" + k]).max() <= 2e-6, k


def test_oracle_f64_matches_reference(seed1_weights, golden):
    logits = O.forward(golden["x"][:6], seed1_weights, np.float64)
    assert np.abs(logits - golden["logits"][:6]).max() <= 2e-6
    assert np.array_equal(O.labels_from_logits(logits), golden["labels"][:6])


def test_oracle_edge_cases(seed1_weights, edge):
    logits = O.forward(edge["x"], seed1_weights, np.float32)
    assert np.abs(logits - edge["logits"]).max() <= 2e-6
    assert np.array_equal(O.labels_from_logits(logits), edge["labels"])


def test_torch_port_is_the_reference_op_sequence(seed1_state, golden, edge):
    port = TorchCpuPort(seed1_state)
    out = port.forward(torch.from_numpy(golden["x"])).numpy()
    assert np.abs(out - golden["logits"]).max() <= 1e-6
    assert np.array_equal(port.predict(torch.from_numpy(edge["x"])).numpy(), edge["labels"])


def test_golden_is_not_degenerate(golden):
    hist = np.bincount(golden["labels"].ravel(), minlength=5)
    assert (hist > 0).sum() >= 3                                         # mixed labels
    assert O.top2_gap(golden["logits"]).min() < 1e-3                     # contains near ties


def test_oracle_rejects_bad_input(seed1_weights):
    with pytest.raises(ValueError):
        O.forward(np.zeros((1, 200, 30), np.uint8), seed1_weights)       # BASELINE.json's shape is not runnable
    x = np.zeros((1, 200, 90), np.uint8)
    x[0, 0, 0] = 12
    with pytest.raises(IndexError):
        O.forward(x, seed1_weights)


def test_empty_batch(seed1_weights):
    assert O.forward(np.zeros((0, 200, 90), np.uint8), seed1_weights).shape == (0, 90, 5)


def test_window_independence(seed1_weights):
    x = structured_windows(3, seed=5)
    full = O.forward(x, seed1_weights)
    one = O.forward(x[1:2], seed1_weights)
    assert np.abs(full[1:2] - one).max() <= 1e-6


def test_synth_generators_are_deterministic():
    a, b = structured_windows(4, seed=9), structured_windows(4, seed=9)
    assert a.dtype == np.uint8 and a.shape == (4, 200, 90) and np.array_equal(a, b)
    assert a.max() <= 11 and (a == 5).any() and (a >= 6).any()
    u = uniform_windows(2, seed=1)
    assert u.max() == 11 and u.min() == 0
