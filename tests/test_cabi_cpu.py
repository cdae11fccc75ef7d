"""CPU: the C-ABI library loads and exports every symbol include/roko_b200.h declares; the host-side
mirror of the reference interface behaves (names, state_dict, error behaviour).  No compute calls."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from roko_b200 import _cabi
    return _cabi.lib()


def header_symbols():
    src = open(os.path.join(ROOT, "include", "roko_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(roko_b200_[a-z0-9_]+)\s*\(", src)))


def test_every_declared_symbol_is_exported_and_bound(lib):
    from roko_b200 import _cabi
    names = header_symbols()
    assert len(names) >= 15
    for n in names:
        assert hasattr(lib, n), n
    assert sorted(_cabi.SIGNATURES) == names            # the binding covers exactly the header


def test_geometry_queries(lib):
    assert lib.roko_b200_abi_version() == 1
    assert (lib.roko_b200_window_reads(), lib.roko_b200_window_cols(), lib.roko_b200_num_classes()) == (200, 90, 5)
    assert lib.roko_b200_raw_weight_count() == 1099731
    per = lib.roko_b200_workspace_bytes(1)
    assert per == (90 * 512 + 90 * 768 + 2 * 90 * 256) * 4 + 18000
    assert lib.roko_b200_workspace_bytes(128) == 128 * per


def test_training_scratch_follows_the_chain_selection(lib, monkeypatch):
    """Default chain: no materialised masked embedding (4.7 MB of saved activations per window); the A/B chains
    ROKO_B200_TRAIN_TC <= 4 keep it (200 x 90 x 50 fp32 more).  Same variable, same reading as model creation."""
    monkeypatch.delenv("ROKO_B200_TRAIN_TC", raising=False)
    per = lib.roko_b200_train_workspace_bytes(1)
    assert 4.5e6 < per < 5.0e6 and lib.roko_b200_train_workspace_bytes(128) == 128 * per
    monkeypatch.setenv("ROKO_B200_TRAIN_TC", "3")
    assert lib.roko_b200_train_workspace_bytes(1) == per + 200 * 90 * 50 * 4
    monkeypatch.setenv("ROKO_B200_TRAIN_TC", "6")
    assert lib.roko_b200_train_workspace_bytes(1) == per


def test_argument_errors_without_gpu(lib):
    from roko_b200 import _cabi
    assert lib.roko_b200_model_create(None, 0) == _cabi.EARG
    assert lib.roko_b200_forward_u8(None, None, 1, None, None, None, 0, None) == _cabi.EARG
    assert b"NULL" in lib.roko_b200_last_error()
    if not torch.cuda.is_available():
        p = _cabi.c_model_p()
        assert lib.roko_b200_model_create(ctypes.byref(p), 0) != _cabi.OK       # fails loudly, no fallback


def test_rnn_mirrors_reference_interface(seed1_state):
    import roko_b200.rnn_model as rm
    for name in ("RNN", "IN_SIZE", "HIDDEN_SIZE", "NUM_LAYERS", "gru_init", "nn", "F", "torch", "np", "math", "init"):
        assert hasattr(rm, name), name                   # callers star-import (inference.py:9, train.py:10)
    assert (rm.IN_SIZE, rm.HIDDEN_SIZE, rm.NUM_LAYERS) == (500, 128, 3)
    m = rm.RNN(rm.IN_SIZE, rm.HIDDEN_SIZE, rm.NUM_LAYERS)
    assert list(m.state_dict().keys()) == list(seed1_state.keys())
    m.load_state_dict(seed1_state, strict=True)
    assert all(torch.equal(v, seed1_state[k]) for k, v in m.state_dict().items())
    assert sum(p.numel() for p in m.parameters()) == 1099731
    with pytest.raises(ValueError):
        rm.RNN(500, 64, 3)


def test_same_seed_gives_reference_init(seed1_state):
    """Module construction order + gru_init replicate the reference's RNG consumption."""
    import roko_b200.rnn_model as rm
    threads = torch.get_num_threads()
    try:
        torch.set_num_threads(1)                         # the fixture was made single-threaded (QR is thread dependent)
        torch.manual_seed(1)
        m = rm.RNN(500, 128, 3)
    finally:
        torch.set_num_threads(threads)
    for k, v in m.state_dict().items():
        assert torch.equal(v, seed1_state[k]), k


def test_no_cpu_fallback(seed1_state, golden):
    import roko_b200.rnn_model as rm
    m = rm.RNN(500, 128, 3).eval()
    m.load_state_dict(seed1_state)
    with torch.no_grad(), pytest.raises(RuntimeError, match="CUDA"):
        m(torch.from_numpy(golden["x"][:1]))
    with pytest.raises(RuntimeError):
        m.predict(torch.zeros((1, 200, 30), dtype=torch.uint8))          # wrong geometry


def test_module_copies_do_not_share_device_handles(seed1_state):
    import copy, pickle
    import roko_b200.rnn_model as rm
    m = rm.RNN(500, 128, 3)
    m.load_state_dict(seed1_state)
    m._handles["fake"] = object()                       # stands in for a live C handle
    for clone in (copy.deepcopy(m), pickle.loads(pickle.dumps(m))):
        assert clone._handles == {}
        assert all(torch.equal(a, b) for a, b in zip(clone.state_dict().values(), m.state_dict().values()))


def test_set_option_rejects_unknown_names(lib):
    from roko_b200 import _cabi
    assert lib.roko_b200_model_set_option(None, b"rec_tc_min", 128) == _cabi.EARG


def test_missing_library_fails_loudly(monkeypatch):
    from roko_b200 import _cabi
    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_cabi, "LIB_PATH", "/nonexistent/libroko_b200.so")
    with pytest.raises(RuntimeError, match="no fallback"):
        _cabi.lib()
