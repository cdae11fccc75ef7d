import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu under gpurun)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session")
def seed1_state():
    import torch
    return torch.load(os.path.join(GOLDEN, "rand_seed1.pth"), map_location="cpu")


@pytest.fixture(scope="session")
def seed1_weights(seed1_state):
    return {k: v.numpy() for k, v in seed1_state.items()}


@pytest.fixture(scope="session")
def golden():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "golden_seed1.npz")))


@pytest.fixture(scope="session")
def edge():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "edge_seed1.npz")))


@pytest.fixture(scope="session")
def cuda_model(seed1_state):
    """The product path: roko_b200.RNN on cuda:0 with the golden weights (fails loudly w/o GPU/.so).

    Inference fixture: parameters frozen, so ``model(x)`` takes the inference kernels whether or not
    the caller wrapped it in ``torch.no_grad()`` (the trainable twin is ``train_model``)."""
    import torch
    from roko_b200.rnn_model import RNN, IN_SIZE, HIDDEN_SIZE, NUM_LAYERS
    m = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS)
    m.load_state_dict(seed1_state, strict=True)
    return m.to("cuda:0").eval().requires_grad_(False)


@pytest.fixture()
def train_model(seed1_state):
    """A fresh trainable roko_b200.RNN on cuda:0 with the golden weights."""
    from roko_b200.rnn_model import RNN, IN_SIZE, HIDDEN_SIZE, NUM_LAYERS
    m = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS)
    m.load_state_dict(seed1_state, strict=True)
    return m.to("cuda:0")


@pytest.fixture(scope="session")
def train_golden():
    import numpy as np
    return dict(np.load(os.path.join(GOLDEN, "train_seed1.npz")))


@pytest.fixture()
def make_model(seed1_state):
    """Factory of fresh inference models on cuda:0 with C-library options applied (kernel A/B selection)."""
    def make(**options):
        from roko_b200.rnn_model import RNN, IN_SIZE, HIDDEN_SIZE, NUM_LAYERS
        m = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS)
        m.load_state_dict(seed1_state, strict=True)
        m = m.to("cuda:0").eval().requires_grad_(False)
        for k, v in options.items():
            m.set_option(k, v)
        return m
    return make


@pytest.fixture(scope="session")
def golden_b128():
    import numpy as np
    from roko_b200.synth import structured_windows
    g = dict(np.load(os.path.join(GOLDEN, "golden_b128_seed1.npz")))
    g["x"] = structured_windows(128, seed=int(g["seed"]))
    assert int(g["x"].astype(np.int64).sum()) == int(g["x_crc"])
    return g
