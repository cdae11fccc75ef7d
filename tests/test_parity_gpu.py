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


# ---- the batch shapes of BASELINE.json configs 2-4 and the tensor-core kernels at those shapes -------------
FFMA = dict(proj=0, rec_tc_min=0, front=0)     # fp32 A/B configuration: FFMA projection + register-resident FFMA recurrence + round-1 front end


def _oracle_subset(logits, labels, x, weights, idx):
    """Windows are independent, so the fp64 oracle only has to evaluate a sample of them."""
    ref64 = O.forward(x[idx], weights, np.float64)
    err = np.abs(logits[idx] - ref64).max()
    gap = O.top2_gap(ref64)
    mism = labels[idx] != O.labels_from_logits(ref64)
    assert err <= TOL_TIGHT, err
    assert not (mism & (gap >= GAP_EXACT)).any()
    return err


def _subset(n, k=48, seed=0):
    rng = np.random.default_rng(seed)
    fixed = [0, 1, 15, 16, 31, 32, 33, n // 2, n - 33, n - 32, n - 17, n - 16, n - 1]
    idx = sorted(set(i for i in fixed if 0 <= i < n) | set(rng.integers(0, n, size=k).tolist()))
    return np.asarray(idx)


@pytest.mark.parametrize("opts", [{}, dict(proj=3, rec=1, front=0)], ids=["fp16", "round1-tf32"])
def test_tensor_core_recurrence_at_batch_128(make_model, golden_b128, opts):
    """The exact shape bench.py times: one 128-window batch, recurrence on tcgen05 (4 CTAs per direction)."""
    m = make_model(rec_tc_min=64, **opts)
    labels, logits = m.predict(_dev(golden_b128["x"]), return_logits=True)
    err = np.abs(logits.cpu().numpy() - golden_b128["logits"]).max()
    print("batch-128 max abs logit err vs reference class", err)
    assert err <= TOL_TIGHT
    assert np.array_equal(labels.cpu().numpy(), golden_b128["labels"])
    m.check_codes()


def test_config3_batch_1024(cuda_model, make_model, seed1_weights):
    """BASELINE.json configs[2]: batch = 1024.  Oracle on a sample of windows; every window against the FFMA kernels."""
    x = structured_windows(1024, seed=1524)
    labels, logits = cuda_model.predict(_dev(x), return_logits=True)
    logits, labels = logits.cpu().numpy(), labels.cpu().numpy()
    err = _oracle_subset(logits, labels, x, seed1_weights, _subset(1024))
    lab2, log2 = make_model(**FFMA).predict(_dev(x), return_logits=True)
    log2 = log2.cpu().numpy()
    d = np.abs(logits - log2).max()
    print(f"B=1024: oracle-sample err {err:.2e}; max |tensor-core - FFMA| {d:.2e}")
    assert d <= TOL_TIGHT
    mism = labels != lab2.cpu().numpy()
    assert not (mism & (O.top2_gap(log2.astype(np.float64)) >= GAP_EXACT)).any()


def test_config2_1000_windows_at_batch_128(cuda_model, seed1_weights):
    """BASELINE.json configs[1]: 1 000 windows in batches of 128 (7 x 128 + 104, inference.py:105) == one call over all
    of them, bit for bit, and == the oracle on a sample."""
    x = structured_windows(1000, seed=1000)
    xd = _dev(x)
    parts = [cuda_model.predict(xd[i:i + 128], return_logits=True) for i in range(0, 1000, 128)]
    labels = torch.cat([p[0] for p in parts])
    logits = torch.cat([p[1] for p in parts])
    whole_labels, whole_logits = cuda_model.predict(xd, return_logits=True)
    assert torch.equal(labels, whole_labels) and torch.equal(logits, whole_logits)
    _oracle_subset(logits.cpu().numpy(), labels.cpu().numpy(), x, seed1_weights, _subset(1000))


def test_full_pass_2368_windows(cuda_model, make_model, seed1_weights):
    """One whole device pass of predict_host (148 SMs x 16 windows): every persistent kernel with all its CTAs."""
    x = np.concatenate([structured_windows(1184, seed=2368), uniform_windows(1184, seed=2369)])
    labels, logits = cuda_model.predict(_dev(x), return_logits=True)
    logits, labels = logits.cpu().numpy(), labels.cpu().numpy()
    _oracle_subset(logits, labels, x, seed1_weights, _subset(2368))
    lab2, log2 = make_model(**FFMA).predict(_dev(x), return_logits=True)
    assert np.abs(logits - log2.cpu().numpy()).max() <= TOL_TIGHT
    host = cuda_model.predict_host(torch.from_numpy(x).pin_memory(), batch=128)
    assert np.array_equal(host.numpy(), labels)
    cuda_model.check_codes()


def test_graph_replay_equals_direct_launches(make_model, golden):
    """The CUDA-graph replay of the chain (default) and plain launches give the same bytes, call after call,
    with different input / output buffers per call."""
    a, b = make_model(graphs=1), make_model(graphs=0)
    xs = [_dev(np.roll(golden["x"], i, axis=0)) for i in range(4)]
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):                      # graphs need a real (capturable) stream
        for x in xs:
            la, ga = a.predict(x, return_logits=True)
            lb, gb = b.predict(x, return_logits=True)
            assert torch.equal(la, lb) and torch.equal(ga, gb)
            assert torch.equal(a.predict(x), la)    # labels-only instance
    s.synchronize()
    assert np.array_equal(la.cpu().numpy(), np.roll(golden["labels"], 3, axis=0))


def test_weights_written_through_data_need_invalidate(make_model, golden):
    """writes through ``param.data`` do not bump the version counter; ``invalidate()`` is the contract."""
    m = make_model()
    x = _dev(golden["x"][:2])
    before = m(x)
    with torch.no_grad():
        m.fc4.bias.add_(1.0)                        # counted: re-packed automatically
    assert torch.allclose(m(x), before + 1.0, atol=1e-6)
    m.fc4.bias.data.sub_(1.0)                       # NOT counted by torch ...
    m.invalidate()                                  # ... so the caller says so
    assert torch.allclose(m(x), before, atol=1e-6)


def test_fp16_range_guard(make_model, golden, seed1_state):
    """A GRU weight beyond the fp16-split range is reported by check_codes(), and the tf32 kernels still serve it."""
    from roko_b200._cabi import RokoB200Error
    sd = {k: v.clone() for k, v in seed1_state.items()}
    sd["gru.weight_hh_l1"][5, 7] = 300.0
    m = make_model()
    m.load_state_dict(sd)
    m.predict(_dev(golden["x"][:1]))
    with pytest.raises(RokoB200Error):
        m.check_codes()


def test_large_front_end_activations(make_model, seed1_state, golden):
    """Trained models have larger fc1 outputs than a random initialisation: with the embedding scaled x80 (fc1 outputs of
    several hundred) the fp16-split front end (operand scales chosen for a < 4 094) still agrees with the fp32-operand round-1 kernels."""
    sd = {k: v.clone() for k, v in seed1_state.items()}
    sd["embedding.weight"] *= 80.0
    a, b = make_model(), make_model(front=0, proj=0, rec_tc_min=0)
    a.load_state_dict(sd)
    b.load_state_dict(sd)
    x = _dev(golden["x"])
    ta, tb = a.forward_taps(x), b.forward_taps(x)
    scale = float(tb["front"].abs().max())
    print("max front activation", scale)
    assert scale > 50.0
    assert float((ta["front"] - tb["front"]).abs().max()) <= 2e-6 * scale
    assert float((ta["logits"] - tb["logits"]).abs().max()) <= TOL
    a.check_codes()
