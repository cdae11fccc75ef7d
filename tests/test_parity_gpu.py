"""GPU parity: the CUDA path (through the C ABI) vs the reference-generated golden vectors and
vs the numpy oracle.  Tolerances: |logit error| <= 1e-4 (north star), target <= 5e-6; labels
bit-exact on every position whose fp64 top-2 logit gap is >= GAP_EXACT."""
import numpy as np
import pytest
import torch

from oracle import roko_oracle as O
from roko_b200.synth import structured_windows, uniform_windows

pytestmark = pytest.mark.gpu

TOL = 1e-4          # north-star tolerance on logits
TOL_TIGHT = 5e-6    # what fp32-accurate kernels should reach
GAP_EXACT = 4e-6    # positions with a smaller fp64 top-2 gap are ambiguous even between fp32 CPU runs


def _dev(x):
    return torch.from_numpy(x).to("cuda:0")


def test_stage_taps_match_reference(cuda_model, golden):
    t = cuda_model.forward_taps(_dev(golden["x"][:2]))
    for k in ("front", "gru_l0", "gru_l1", "gru_l2"):
        err = np.abs(t[k].cpu().numpy() - golden["tap_" + k]).max()
        print(k, "max abs err", err)
        assert err <= TOL_TIGHT, (k, err)


def test_golden_logits_and_labels(cuda_model, golden):
    x = _dev(golden["x"])
    logits = cuda_model(x).cpu().numpy()
    labels = cuda_model.predict(x).cpu().numpy()
    err = np.abs(logits - golden["logits"]).max()
    print("golden max abs logit err", err)
    assert err <= TOL and err <= TOL_TIGHT
    assert np.array_equal(labels, golden["labels"])
    assert np.array_equal(labels, O.labels_from_logits(logits))


def test_edge_cases(cuda_model, edge):
    x = _dev(edge["x"])
    logits = cuda_model(x).cpu().numpy()
    assert np.abs(logits - edge["logits"]).max() <= TOL_TIGHT
    assert np.array_equal(cuda_model.predict(x).cpu().numpy(), edge["labels"])


def test_int64_input_matches_uint8(cuda_model, golden):
    x = _dev(golden["x"][:5])
    a = cuda_model(x)
    b = cuda_model(x.to(torch.int64))          # what the reference caller passes (inference.py:113)
    assert torch.equal(a, b)


@pytest.mark.parametrize("batch", [1, 7, 104, 128, 300])
def test_oracle_parity_batches(cuda_model, seed1_weights, batch):
    """Ragged batch sizes incl. the 104-window tail of 1000 windows @128 (inference.py:105)."""
    x = structured_windows(batch, seed=500 + batch)
    ref64 = O.forward(x, seed1_weights, np.float64)
    labels, logits = cuda_model.predict(_dev(x), return_logits=True)
    logits, labels = logits.cpu().numpy(), labels.cpu().numpy()
    err = np.abs(logits - ref64).max()
    gap = O.top2_gap(ref64)
    ref_labels = O.labels_from_logits(ref64)
    mism = labels != ref_labels
    print(f"B={batch}: max err {err:.2e}; label hist {np.bincount(labels.ravel(), minlength=5)}; "
          f"gaps<1e-4: {(gap < 1e-4).sum()}, min gap {gap.min():.2e}; mismatches {mism.sum()}")
    assert err <= TOL and err <= TOL_TIGHT
    assert not (mism & (gap >= GAP_EXACT)).any()
    assert mism.sum() == 0 or gap[mism].max() < GAP_EXACT


def test_chunked_equals_unchunked(cuda_model, seed1_weights):
    """A batch larger than the internal chunk gives the same bytes as window-at-a-time."""
    import roko_b200.rnn_model as rm
    x = _dev(uniform_windows(37, seed=9))
    full = cuda_model(x)
    old = rm.MAX_CHUNK
    try:
        rm.MAX_CHUNK = 8
        for h in cuda_model._handles.values():
            h.workspaces.clear()
        chunked = cuda_model(x)
    finally:
        rm.MAX_CHUNK = old
        for h in cuda_model._handles.values():
            h.workspaces.clear()
    assert torch.equal(full, chunked)


def test_batch_invariance(cuda_model):
    """Windows are independent: a window's logits do not depend on its batch neighbours."""
    x = _dev(structured_windows(9, seed=77))
    full = cuda_model(x)
    for i in (0, 4, 8):
        assert torch.equal(full[i:i + 1], cuda_model(x[i:i + 1]))


def test_predict_host_pipeline(cuda_model, golden):
    x = torch.from_numpy(np.concatenate([golden["x"]] * 3)).pin_memory()      # 48 windows, ragged vs batch 20
    out = cuda_model.predict_host(x, batch=20)
    assert np.array_equal(out.numpy(), np.concatenate([golden["labels"]] * 3))


def test_out_of_range_code_raises(cuda_model, golden):
    x = golden["x"][:1].copy()
    x[0, 3, 5] = 12
    cuda_model.predict(_dev(x))
    with pytest.raises(IndexError):
        cuda_model.check_codes()
    cuda_model.predict(_dev(golden["x"][:1]))
    cuda_model.check_codes()                   # flag cleared, clean input passes


def test_cpu_tensor_fails_loudly(cuda_model, golden):
    with pytest.raises(RuntimeError):
        cuda_model(torch.from_numpy(golden["x"][:1]))
