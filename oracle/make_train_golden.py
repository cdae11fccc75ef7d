"""Generate tests/golden/train_seed1.npz by EXECUTING the reference class (build container only).

    python oracle/make_train_golden.py      # needs /root/reference (absent on the GPU box)

Imports /root/reference/roko/rnn_model.py unmodified, loads tests/golden/rand_seed1.pth, and for
3 seeded structured windows with their truth rows as targets computes what the reference's
training step computes before the optimiser update (roko/train.py:46-53), with dropout off
(``model.eval()``: the dropout draws of torch's generator are not reproducible by another
implementation; the dropout sites are checked separately against exported masks):

    logits = model(x);  loss = F.cross_entropy(logits.transpose(1, 2), y);  loss.backward()

The input seed is one whose ReLU pre-activations all stay >= 2e-6 away from 0 in float64, so that fp32
implementations agree with the reference on every ReLU derivative (train_oracle.relu_margin).

A second fixture, train_b128_seed1.npz, holds the same quantities for one full training batch of 128 windows
(x = structured_windows(128, seed=9128), regenerated from the seed; logits of the first 4 windows only): the batch
size of BASELINE.json config 5, where the streamed tensor-core products and split reductions of the CUDA path
run with all their CTAs.  With 57 M ReLU pre-activations some inevitably sit within fp32 noise of 0, so tests
compare it with a looser tolerance (see tests/test_train_gpu.py).

Stored: x, y, logits, loss, and per parameter the gradient's L2 norm, its sum, and its values at
``train_oracle.sample_index`` strided positions (all 31 tensors; small ones in full).
"""
import os
import sys

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/roko")

import rnn_model as ref  # noqa: E402  (the reference, unmodified)
from oracle.train_oracle import sample_index  # noqa: E402
from roko_b200.synth import structured_windows  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def reference_step(x, y):
    model = ref.RNN(ref.IN_SIZE, ref.HIDDEN_SIZE, ref.NUM_LAYERS)
    model.load_state_dict(torch.load(os.path.join(OUT, "rand_seed1.pth")))
    model.eval()
    logits = model(torch.from_numpy(x).type(torch.LongTensor))
    loss = F.cross_entropy(logits.transpose(1, 2), torch.from_numpy(y))
    loss.backward()
    out = {"loss": np.float64(loss.item())}
    for name, p in model.named_parameters():
        g = p.grad.numpy().reshape(-1).astype(np.float64)
        idx = sample_index(g.size)
        out[f"norm/{name}"] = np.float64(np.sqrt((g * g).sum()))
        out[f"sum/{name}"] = np.float64(g.sum())
        out[f"sample/{name}"] = g[idx].astype(np.float32)
    return logits.detach().numpy(), out


def main():
    torch.set_num_threads(1)
    x, truth = structured_windows(3, seed=472, return_truth=True)   # ReLU margin 2.3e-6, see train_oracle.relu_margin
    y = truth.astype(np.int64)
    logits, out = reference_step(x, y)
    out.update({"x": x, "y": y.astype(np.uint8), "logits": logits})
    np.savez_compressed(os.path.join(OUT, "train_seed1.npz"), **out)
    print("loss", out["loss"], "wrote", os.path.join(OUT, "train_seed1.npz"))

    torch.set_num_threads(8)
    xb, tb = structured_windows(128, seed=9128, return_truth=True)
    logits, out = reference_step(xb, tb.astype(np.int64))
    out.update({"seed": np.int64(9128), "logits4": logits[:4]})
    np.savez_compressed(os.path.join(OUT, "train_b128_seed1.npz"), **out)
    print("b128 loss", out["loss"], "wrote", os.path.join(OUT, "train_b128_seed1.npz"))


if __name__ == "__main__":
    main()
