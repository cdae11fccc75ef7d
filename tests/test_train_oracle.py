"""CPU: the training oracle (oracle/train_oracle.py) against gradients of the reference class itself
(tests/golden/train_seed1.npz, made by oracle/make_train_golden.py), and its helpers."""
import numpy as np

from oracle import train_oracle as TO
from roko_b200.synth import structured_windows


def test_oracle_gradients_match_reference_fixture(seed1_weights, train_golden):
    g = train_golden
    logits, loss, grads = TO.loss_and_grads(seed1_weights, g["x"], g["y"])
    assert np.abs(logits - g["logits"]).max() <= 5e-6
    assert abs(loss - float(g["loss"])) <= 1e-6
    for k in TO.STATE_KEYS:
        flat = grads[k].reshape(-1)
        ref = g[f"sample/{k}"].astype(np.float64)
        assert np.abs(flat[TO.sample_index(flat.size)] - ref).max() <= 2e-5 * np.abs(ref).max(), k
        norm = float(g[f"norm/{k}"])
        assert abs(np.sqrt((flat * flat).sum()) - norm) <= 1e-5 * norm, k


def test_fixture_inputs_clear_the_relu_kinks(seed1_weights, train_golden):
    assert TO.relu_margin(seed1_weights, train_golden["x"]) >= 2e-6
    x = structured_windows(1, seed=7126)
    assert TO.relu_margin(seed1_weights, x) >= 1.5e-6


def test_mask_restatement_statistics():
    a = TO.kernel_keep_masks(0.2, 11, 2)
    b = TO.kernel_keep_masks(0.2, 12, 2)
    shapes = TO.mask_shapes(2)
    for k in TO.SITES:
        assert a[k].shape == shapes[k] and a[k].dtype == np.uint8
        assert abs(a[k].mean() - 0.8) < 0.01
        assert (a[k] != b[k]).mean() > 0.2
    assert not np.array_equal(a["gru0"], a["gru1"])
    assert all(v.all() for v in TO.kernel_keep_masks(0.0, 11, 1).values())
    # masks are a function of (seed, site, element index): a bigger batch extends a smaller one
    c = TO.kernel_keep_masks(0.2, 11, 1)
    assert all(np.array_equal(c[k][0], a[k][0]) for k in TO.SITES)


def test_mask_restatement_known_answers():
    """Pins the restatement of roko_b200/csrc/train.cuh:drop_hash (the GPU suite pins the kernels to the restatement)."""
    m = TO.kernel_keep_masks(0.2, 20240921, 1)
    head = {k: "".join(str(int(v)) for v in m[k].reshape(-1)[:48]) for k in ("emb", "fc1", "gru1")}
    assert head == {"emb": "111011111111111101101111101011011110101011111111",
                    "fc1": "110101001101110110111111110110101110111111011111",
                    "gru1": "111111100101111111111101001111111111111101101111"}
    assert (int(m["emb"].sum()), int(m["fc1"].sum()), int(m["gru1"].sum())) == (720556, 359650, 18406)


def test_masked_forward_differs_and_scales(seed1_weights, train_golden):
    x, y = train_golden["x"][:1], train_golden["y"][:1]
    masks = TO.kernel_keep_masks(0.2, 3, 1)
    l0, _, _ = TO.loss_and_grads(seed1_weights, x, y)
    l1, _, g1 = TO.loss_and_grads(seed1_weights, x, y, masks, 0.2)
    assert np.abs(l0 - l1).max() > 1e-3
    assert all(np.isfinite(v).all() for v in g1.values())
