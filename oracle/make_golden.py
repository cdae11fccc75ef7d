"""Generate tests/golden/* by EXECUTING the reference class (run in the build container only).

    python oracle/make_golden.py            # needs /root/reference (absent on the GPU box)

Imports /root/reference/roko/rnn_model.py unmodified, builds ``RNN(500,128,3)`` under
``torch.manual_seed(1)`` (mixed labels; seed 0 predicts one class -- SURVEY.md section 8d),
runs it on seeded structured pileups on the CPU in eval mode and stores

  rand_seed1.pth          the reference module's own state_dict (the .pth contract, App. A)
  golden_seed1.npz        x (16,200,90) u8, logits (16,90,5) f32, labels (16,90) u8 (argmax as
                          inference.py:116), stage taps for the first 2 windows
                          (front (2,90,500), gru_l0..2 (2,90,256))
  golden_b128_seed1.npz   one full inference batch: x = structured_windows(128, seed=501) (not stored: regenerated
                          from the seed), logits (128,90,5) f32, labels (128,90) u8 -- the batch of BASELINE.json
                          configs[1]; bench.py's parity gate and the tensor-core-recurrence tests run on it
  edge_seed1.npz          edge-case windows (all one code, all UNKNOWN, codes 0 and 11 only,
                          single strand) with logits/labels

This file is the provenance of the fixtures; nothing under tests/ or bench.py reads
/root/reference at run time.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference/roko")

import rnn_model as ref  # noqa: E402  (the reference, unmodified)
from roko_b200.synth import structured_windows, uniform_windows  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def run_ref(model, x_u8, taps=False):
    x = torch.from_numpy(x_u8).type(torch.LongTensor)              # inference.py:113
    with torch.no_grad():
        if not taps:
            logits = model(x)
            return logits.numpy(), torch.argmax(logits, dim=2).numpy().astype(np.uint8), {}
        # re-trace forward() stage by stage to export taps (same ops, same order)
        t = {}
        h = model.embedding(x).permute((0, 2, 3, 1))
        h = torch.relu(model.fc1(h))
        h = torch.relu(model.fc2(h))
        h = h.reshape(-1, 90, ref.IN_SIZE)
        t["front"] = h.numpy().copy()
        # per-layer taps: run the stacked GRU layer by layer with its own weights
        v = h
        for l in range(ref.NUM_LAYERS):
            g = torch.nn.GRU(v.shape[2], ref.HIDDEN_SIZE, num_layers=1, batch_first=True, bidirectional=True)
            for sfx in ("", "_reverse"):
                for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh"):
                    getattr(g, f"{kind}_l0{sfx}").data.copy_(getattr(model.gru, f"{kind}_l{l}{sfx}").data)
            v, _ = g(v)
            t[f"gru_l{l}"] = v.numpy().copy()
        logits = model(x)
        assert torch.equal(model.fc4(v), logits) or (model.fc4(v) - logits).abs().max() < 1e-6
        return logits.numpy(), torch.argmax(logits, dim=2).numpy().astype(np.uint8), t


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(1)                      # bit-reproducible regardless of host cores
    torch.manual_seed(1)
    model = ref.RNN(ref.IN_SIZE, ref.HIDDEN_SIZE, ref.NUM_LAYERS).eval()
    torch.save(model.state_dict(), os.path.join(OUT, "rand_seed1.pth"))

    x = structured_windows(16, seed=101)
    logits, labels, _ = run_ref(model, x)
    _, _, taps = run_ref(model, x[:2], taps=True)
    np.savez_compressed(os.path.join(OUT, "golden_seed1.npz"), x=x, logits=logits, labels=labels,
                        **{f"tap_{k}": v for k, v in taps.items()})
    print("golden: label hist", np.bincount(labels.ravel(), minlength=5),
          "min top-2 gap", float(np.min(np.sort(logits, 2)[..., -1] - np.sort(logits, 2)[..., -2])))

    xb = structured_windows(128, seed=501)
    lb, yb, _ = run_ref(model, xb)
    np.savez_compressed(os.path.join(OUT, "golden_b128_seed1.npz"), seed=np.int64(501), logits=lb, labels=yb,
                        x_crc=np.int64(int(np.frombuffer(xb.tobytes(), dtype=np.uint8).astype(np.int64).sum())))
    print("b128: label hist", np.bincount(yb.ravel(), minlength=5))

    edge = np.zeros((6, 200, 90), dtype=np.uint8)
    edge[0][:] = 0                                 # all 'A' forward
    edge[1][:] = 5                                 # all UNKNOWN forward
    edge[2][:] = 11                                # all UNKNOWN reverse (max code)
    edge[3] = np.where(uniform_windows(1, 7)[0] % 2 == 0, 0, 11)   # only codes 0 and 11
    edge[4] = structured_windows(1, seed=202)[0] % 6               # single (forward) strand
    edge[5] = uniform_windows(1, 8)[0]                             # uniform random codes
    el, ey, _ = run_ref(model, edge)
    np.savez_compressed(os.path.join(OUT, "edge_seed1.npz"), x=edge, logits=el, labels=ey)
    print("edge: label hist", np.bincount(ey.ravel(), minlength=5))


if __name__ == "__main__":
    main()
