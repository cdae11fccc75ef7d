"""GPU parity of the training path: train-mode forward with dropout and
the hand-written backward, through the C ABI behind torch.autograd, against
  * gradients of the reference class itself (tests/golden/train_seed1.npz, dropout off), and
  * the float64 autograd oracle (oracle/train_oracle.py) fed the kernels' own dropout masks.
Tolerance: every gradient tensor within GRAD_TOL of its own max magnitude (fp32 kernels, sums over
up to 576 000 rows; measured <= 2e-5), logits within 5e-6."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import train_oracle as TO
from roko_b200 import rnn_model as RM
from roko_b200.synth import structured_windows

pytestmark = pytest.mark.gpu

GRAD_TOL = 1e-4
LOGIT_TOL = 5e-6
RELU_MARGIN = 1.5e-6   # ReLU pre-activations of the test inputs clear fp32 noise (~3e-7): see train_oracle.relu_margin

# Inputs found offline (oracle only) whose float64 ReLU pre-activations all keep RELU_MARGIN from 0; nearer
# ones make any two implementations disagree on a ReLU derivative, which moves a few entries by ~1e-3.
CLEAR_SEEDS = {1: 7126, 2: 7217, 5: 7515}                       # batch -> structured_windows seed, dropout off
CLEAR_DROPOUT = {2: (8214, 17328), 3: (11980, 22123)}           # batch -> (input seed, mask seed), p = 0.2


def _loss(model, x, y, seed=None):
    xt = torch.from_numpy(x).to("cuda:0")
    yt = torch.from_numpy(y.astype(np.int64)).to("cuda:0")
    logits = model(xt) if seed is None else model._train_forward(xt, seed=seed)
    return logits, F.cross_entropy(logits.transpose(1, 2), yt)      # roko/train.py:49-52


def _grads(model):
    return {k: p.grad.detach().cpu().numpy().astype(np.float64) for k, p in model.named_parameters()}


def _compare(got, want, tol=GRAD_TOL):
    worst = ("", 0.0)
    for k in TO.STATE_KEYS:
        scale = np.abs(want[k]).max() + 1e-30
        err = np.abs(got[k] - want[k]).max() / scale
        if err > worst[1]:
            worst = (k, err)
        assert np.isfinite(got[k]).all(), k
        assert err <= tol, (k, err)
    print("worst relative gradient error", worst)


def test_gradients_match_reference_fixture(train_model, train_golden):
    """Dropout off (eval mode, autograd on): logits, loss and all 31 gradients vs the reference's."""
    g = train_golden
    train_model.eval()
    logits, loss = _loss(train_model, g["x"], g["y"])
    loss.backward()
    assert np.abs(logits.detach().cpu().numpy() - g["logits"]).max() <= LOGIT_TOL
    assert abs(loss.item() - float(g["loss"])) <= 1e-5
    got = _grads(train_model)
    for k in TO.STATE_KEYS:
        flat = got[k].reshape(-1)
        ref = g[f"sample/{k}"].astype(np.float64)
        scale = np.abs(ref).max() + 1e-30
        assert np.abs(flat[TO.sample_index(flat.size)] - ref).max() / scale <= GRAD_TOL, k
        norm = float(g[f"norm/{k}"])
        assert abs(np.sqrt((flat * flat).sum()) - norm) / norm <= GRAD_TOL, k
        assert abs(flat.sum() - float(g[f"sum/{k}"])) <= GRAD_TOL * norm * np.sqrt(flat.size), k


def test_gradients_at_batch_128_match_reference_fixture(train_model):
    """The timed training shape (BASELINE.json config 5: 128 windows per GPU): the streamed tcgen05 products split
    576 000 rows over 148 CTAs and the generic GEMM accumulates with atomics -- all of that against gradients of
    the reference class itself (tests/golden/train_b128_seed1.npz).  Tolerance 1e-3 of each tensor's max, not 1e-4:
    among 57 M ReLU pre-activations some lie within fp32 noise of 0, where any two fp32 implementations pick
    different sides of the kink (the reference in fp32 vs fp64 does too); each such flip moves a few entries."""
    import os
    g = dict(np.load(os.path.join(os.path.dirname(__file__), "golden", "train_b128_seed1.npz")))
    x, y = structured_windows(128, seed=int(g["seed"]), return_truth=True)
    train_model.eval()
    logits, loss = _loss(train_model, x, y)
    loss.backward()
    assert np.abs(logits.detach().cpu().numpy()[:4] - g["logits4"]).max() <= LOGIT_TOL
    assert abs(loss.item() - float(g["loss"])) <= 1e-5
    got = _grads(train_model)
    worst = ("", 0.0)
    for k in TO.STATE_KEYS:
        flat = got[k].reshape(-1)
        assert np.isfinite(flat).all(), k
        ref = g[f"sample/{k}"].astype(np.float64)
        err = np.abs(flat[TO.sample_index(flat.size)] - ref).max() / (np.abs(ref).max() + 1e-30)
        worst = max(worst, (k, err), key=lambda t: t[1])
        assert err <= 1e-3, (k, err)
        norm = float(g[f"norm/{k}"])
        assert abs(np.sqrt((flat * flat).sum()) - norm) / norm <= 1e-3, k
    print("batch-128 worst relative gradient error", worst)


@pytest.mark.parametrize("batch", [1, 2, 5])
def test_gradients_match_oracle_no_dropout(train_model, seed1_weights, batch):
    x, y = structured_windows(batch, seed=CLEAR_SEEDS[batch], return_truth=True)
    assert TO.relu_margin(seed1_weights, x) >= RELU_MARGIN
    train_model.eval()
    logits, loss = _loss(train_model, x, y)
    loss.backward()
    ref_logits, ref_loss, ref = TO.loss_and_grads(seed1_weights, x, y)
    assert np.abs(logits.detach().cpu().numpy() - ref_logits).max() <= LOGIT_TOL
    assert abs(loss.item() - ref_loss) <= 1e-5
    _compare(_grads(train_model), ref)


@pytest.mark.parametrize("batch", [2, 3])
def test_dropout_forward_backward_match_oracle_with_same_masks(train_model, seed1_weights, batch):
    """Train mode, p = 0.2: export the kernels' masks, replay them in the float64 oracle."""
    xseed, seed = CLEAR_DROPOUT[batch]
    x, y = structured_windows(batch, seed=xseed, return_truth=True)
    masks = {k: v.cpu().numpy() for k, v in RM.dropout_masks(0.2, seed, batch, "cuda:0").items()}
    want = TO.kernel_keep_masks(0.2, seed, batch)
    assert all(np.array_equal(masks[k], want[k]) for k in want)      # device masks == their numpy restatement
    assert TO.relu_margin(seed1_weights, x, masks, 0.2) >= RELU_MARGIN
    train_model.train()
    logits, loss = _loss(train_model, x, y, seed=seed)
    loss.backward()
    ref_logits, ref_loss, ref = TO.loss_and_grads(seed1_weights, x, y, masks, 0.2)
    assert np.abs(logits.detach().cpu().numpy() - ref_logits).max() <= 2e-5
    assert abs(loss.item() - ref_loss) <= 1e-5
    _compare(_grads(train_model), ref)


def test_large_seed_masks_match_restatement():
    seed = 2 ** 61 + 7
    got = RM.dropout_masks(0.2, seed, 1, "cuda:0")
    want = TO.kernel_keep_masks(0.2, seed, 1)
    assert all(np.array_equal(got[k].cpu().numpy(), want[k]) for k in want)


def test_dropout_mask_statistics():
    a = RM.dropout_masks(0.2, 1, 4, "cuda:0")
    b = RM.dropout_masks(0.2, 2, 4, "cuda:0")
    for k in a:
        keep = a[k].float().mean().item()
        assert abs(keep - 0.8) < 0.01, (k, keep)
        assert (a[k] != b[k]).float().mean().item() > 0.2, k       # another seed, another mask
        assert torch.equal(a[k], RM.dropout_masks(0.2, 1, 4, "cuda:0")[k])
    assert not torch.equal(a["gru0"], a["gru1"])                    # sites are independent
    ones = RM.dropout_masks(0.0, 1, 1, "cuda:0")
    assert all(bool(v.all()) for v in ones.values())


def test_train_mode_draws_from_torch_generator(train_model, train_golden):
    x = torch.from_numpy(train_golden["x"]).to("cuda:0")
    train_model.train()
    torch.manual_seed(5)
    a = train_model(x)
    torch.manual_seed(5)
    b = train_model(x)
    c = train_model(x)
    assert torch.equal(a, b) and not torch.equal(a, c)
    train_model.eval()
    with torch.no_grad():
        d = train_model(x)
    assert not torch.equal(a, d)
    assert np.abs(d.cpu().numpy() - train_golden["logits"]).max() <= LOGIT_TOL


def test_adam_steps_reduce_the_loss(train_model):
    """The reference's optimiser and loss (roko/train.py:39,52) around the custom forward."""
    x, y = structured_windows(8, seed=901, return_truth=True)
    opt = torch.optim.Adam(train_model.parameters(), lr=1e-3)
    train_model.train()
    torch.manual_seed(0)
    first = last = None
    for _ in range(12):
        opt.zero_grad()
        _, loss = _loss(train_model, x, y)
        loss.backward()
        opt.step()
        first = loss.item() if first is None else first
        last = loss.item()
    print("loss", first, "->", last)
    assert last < 0.8 * first
    train_model.eval()
    with torch.no_grad():                                            # the packed-weight cache followed the updates
        logits = train_model(torch.from_numpy(x).to("cuda:0"))
    assert F.cross_entropy(logits.transpose(1, 2), torch.from_numpy(y.astype(np.int64)).to("cuda:0")).item() < first


def test_training_errors(train_model, train_golden):
    x = torch.from_numpy(train_golden["x"]).to("cuda:0")
    train_model.train()
    out = train_model(x)
    out.sum().backward()
    with pytest.raises(RuntimeError):
        out.sum().backward()                                         # saved activations are consumed
    with pytest.raises(RuntimeError, match="at most"):
        train_model(torch.zeros((RM.MAX_TRAIN_BATCH + 1, 200, 90), dtype=torch.uint8, device="cuda:0"))
    bad = x.clone()
    bad[0, 0, 0] = 12
    train_model(bad)
    with pytest.raises(IndexError):
        train_model.check_codes()


def test_train_driver_end_to_end(tmp_path):
    """roko_b200.train.train (the reference's train.py entry point) on the real model: labelled windows
    from an in-memory .hdf5 stand-in, two epochs, checkpoint interchangeable with the reference's."""
    from roko_b200 import train as T
    from tests import fake_h5
    xs, ys = structured_windows(96, seed=33, return_truth=True)
    pos = np.zeros((96, 90, 2), np.int64)
    fake_h5.register("mem://gpu_train", {"c": "ACGT"}, [("c_0", "c", pos[:64], xs[:64], ys[:64])])
    fake_h5.register("mem://gpu_val", {"c": "ACGT"}, [("c_1", "c", pos[64:], xs[64:], ys[64:])])
    logs = []
    hist = T.train("mem://gpu_train", str(tmp_path), "mem://gpu_val", mem=True, batch_size=16, epochs=3, lr=2e-3,
                   device="cuda:0", h5=fake_h5, log=logs.append, seed=4)
    print(hist)
    assert hist["epochs"] == 3 and hist["checkpoint"]
    assert hist["train_loss"][-1] < hist["train_loss"][0]
    assert hist["val_loss"][-1] < 1.6                                   # below ln 5: it learned something
    sd = torch.load(hist["checkpoint"])
    assert list(sd) == RM.state_keys()                                   # the reference's state_dict contract
    m = RM.RNN(RM.IN_SIZE, RM.HIDDEN_SIZE, RM.NUM_LAYERS)
    m.load_state_dict(sd, strict=True)
