"""Diagnostic for the training path: per-tensor gradient errors vs the float64 oracle (no asserts),
with and without dropout, plus stage timings.   python scripts/train_check.py [batch]"""
import os
import sys
import time

import numpy as np
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import train_oracle as TO  # noqa: E402
from roko_b200 import rnn_model as RM  # noqa: E402
from roko_b200.synth import structured_windows, uniform_windows  # noqa: E402


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    sd = torch.load(os.path.join(ROOT, "tests", "golden", "rand_seed1.pth"))
    state = {k: v.numpy() for k, v in sd.items()}
    m = RM.RNN(500, 128, 3)
    m.load_state_dict(sd)
    m = m.to("cuda:0")
    x, y = structured_windows(batch, seed=int(os.environ.get("TRAIN_SEED", "42")), return_truth=True)
    xt = torch.from_numpy(x).cuda()
    yt = torch.from_numpy(y.astype(np.int64)).cuda()
    for p, seed in ((0.0, 1), (0.2, 777)):
        m.train(p > 0)
        m.zero_grad()
        logits = m._train_forward(xt, seed=seed)
        loss = F.cross_entropy(logits.transpose(1, 2), yt)
        loss.backward()
        torch.cuda.synchronize()
        masks = {k: v.cpu().numpy() for k, v in RM.dropout_masks(p, seed, batch, "cuda:0").items()} if p > 0 else None
        rl, rloss, rg = TO.loss_and_grads(state, x, y, masks, p if p > 0 else 0.2)
        print(f"p={p}: loss {loss.item():.7f} oracle {rloss:.7f}  logits err {np.abs(logits.detach().cpu().numpy() - rl).max():.3e}")
        for k, q in m.named_parameters():
            g = q.grad.detach().cpu().numpy().astype(np.float64)
            scale = np.abs(rg[k]).max() + 1e-30
            err = np.abs(g - rg[k])
            flag = ""
            if err.max() / scale > 2e-5:
                bad = np.argwhere(err > 2e-5 * scale)
                flag = f"  BAD at {bad[:6].tolist()} ({len(bad)} of {g.size}) got {g[tuple(bad[0])]:.6e} want {rg[k][tuple(bad[0])]:.6e}"
            print(f"   {k:30s} max|g| {scale:.3e}  rel err {err.max() / scale:.3e}  finite {np.isfinite(g).all()}{flag}")
    # timing at a training-size batch
    nb = int(os.environ.get("TRAIN_BATCH", "128"))
    if nb <= 0:
        return
    xb = torch.from_numpy(uniform_windows(nb, seed=3)).cuda()
    yb = torch.randint(0, 5, (nb, 90), device="cuda")
    m.train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-4)
    for it in range(6):
        if it == 2:
            torch.cuda.synchronize()
            t0 = time.perf_counter()
        opt.zero_grad()
        loss = F.cross_entropy(m(xb).transpose(1, 2), yb)
        loss.backward()
        opt.step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / 4
    print(f"train step batch {nb}: {dt * 1e3:.2f} ms  -> {nb / dt:.0f} windows/s   peak mem {torch.cuda.max_memory_allocated() / 2**30:.2f} GiB")


if __name__ == "__main__":
    main()
