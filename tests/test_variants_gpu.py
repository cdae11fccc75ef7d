"""GPU: the alternative kernels kept for A/B (selected by environment variables read at model creation)
must satisfy the same parity bar as the defaults.  Each variant runs in a fresh process."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

CHECK = r"""
import numpy as np, torch, sys
sys.path.insert(0, %r)
from roko_b200.rnn_model import RNN, IN_SIZE, HIDDEN_SIZE, NUM_LAYERS
g = np.load(%r)
m = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS)
m.load_state_dict(torch.load(%r, map_location="cpu"))
m = m.to("cuda:0").eval()
x = torch.from_numpy(np.concatenate([g["x"]] * 20)).cuda()            # 320 windows: above the tensor-core recurrence threshold
with torch.no_grad():
    lab, logit = m.predict(x, return_logits=True)
err = float(np.abs(logit.cpu().numpy()[:16] - g["logits"]).max())
same = bool(np.array_equal(lab.cpu().numpy(), np.concatenate([g["labels"]] * 20)))
print("ERR", err, same)
assert err <= 5e-6 and same
""" % (ROOT, os.path.join(ROOT, "tests", "golden", "golden_seed1.npz"), os.path.join(ROOT, "tests", "golden", "rand_seed1.pth"))


@pytest.mark.parametrize("env", [
    {"ROKO_B200_PROJ": "tf32"}, {"ROKO_B200_REC": "tf32"}, {"ROKO_B200_PROJ": "tf32", "ROKO_B200_REC": "tf32"},
    {"ROKO_B200_PROJ": "ffma"}, {"ROKO_B200_REC_TC_MIN": "0"}, {"ROKO_B200_REC_NB": "4"}, {"ROKO_B200_GRAPHS": "0"}, {"ROKO_B200_FRONT": "mma"},
], ids=lambda e: ",".join(f"{k.replace('ROKO_B200_', '')}={v}" for k, v in e.items()))
def test_kernel_variant(env):
    p = subprocess.run([sys.executable, "-c", CHECK], env={**os.environ, **env}, capture_output=True, text=True, timeout=180)
    assert p.returncode == 0, p.stdout + p.stderr


TRAIN_CHECK = r"""
import numpy as np, torch, sys
import torch.nn.functional as F
sys.path.insert(0, %r)
from roko_b200.rnn_model import RNN, IN_SIZE, HIDDEN_SIZE, NUM_LAYERS
from oracle import train_oracle as TO
g = np.load(%r)
m = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS)
m.load_state_dict(torch.load(%r, map_location="cpu"))
m = m.to("cuda:0").eval()
x = torch.from_numpy(g["x"]).cuda()
y = torch.from_numpy(g["y"].astype(np.int64)).cuda()
logits = m(x)
F.cross_entropy(logits.transpose(1, 2), y).backward()
assert float(np.abs(logits.detach().cpu().numpy() - g["logits"]).max()) <= 5e-6
for k, p in m.named_parameters():
    flat = p.grad.detach().cpu().numpy().reshape(-1).astype(np.float64)
    ref = g["sample/" + k].astype(np.float64)
    err = np.abs(flat[TO.sample_index(flat.size)] - ref).max() / (np.abs(ref).max() + 1e-30)
    assert err <= 1e-4, (k, err)
print("OK")
""" % (ROOT, os.path.join(ROOT, "tests", "golden", "train_seed1.npz"), os.path.join(ROOT, "tests", "golden", "rand_seed1.pth"))


@pytest.mark.parametrize("env", [
    {"ROKO_B200_REC_NB": "2"}, {"ROKO_B200_REC_NB": "4"}, {"ROKO_B200_PROJ": "ffma"}, {"ROKO_B200_PROJ": "tf32"},
    {"ROKO_B200_TRAIN_TC": "6"}, {"ROKO_B200_TRAIN_TC": "4"}, {"ROKO_B200_TRAIN_TC": "3"}, {"ROKO_B200_TRAIN_TC": "2"}, {"ROKO_B200_TRAIN_TC": "1"}, {"ROKO_B200_TRAIN_TC": "0"}, {"ROKO_B200_TRAIN_TC": "0", "ROKO_B200_GEMM_NOSTREAM": "1"},
], ids=lambda e: ",".join(f"{k.replace('ROKO_B200_', '')}={v}" for k, v in e.items()))
def test_training_kernel_variant(env):
    """The recurrence's window-group sizes (forward with saved gates, and backward) and the projection
    variants under the training path: gradients vs the reference fixture."""
    p = subprocess.run([sys.executable, "-c", TRAIN_CHECK], env={**os.environ, **env}, capture_output=True, text=True, timeout=180)
    assert p.returncode == 0, p.stdout + p.stderr
