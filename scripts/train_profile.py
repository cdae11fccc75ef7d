"""Forward / backward timing of one training step (CUDA events) at a given batch.
   python scripts/train_profile.py [batch] [iters]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roko_b200 import rnn_model as RM  # noqa: E402
from roko_b200.synth import uniform_windows  # noqa: E402

batch = int(sys.argv[1]) if len(sys.argv) > 1 else 128
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 5
torch.manual_seed(0)
m = RM.RNN(500, 128, 3).to("cuda:0").train()
opt = torch.optim.Adam(m.parameters(), lr=1e-4, fused=True)
x = torch.from_numpy(uniform_windows(batch, seed=3)).cuda()
y = torch.randint(0, 5, (batch, 90), device="cuda")
ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
tf = tb = to = 0.0
for it in range(iters + 2):
    opt.zero_grad()
    ev[0].record()
    logits = m(x)
    ev[1].record()
    loss = F.cross_entropy(logits.transpose(1, 2), y)
    loss.backward()
    ev[2].record()
    opt.step()
    ev[3].record()
    torch.cuda.synchronize()
    if it >= 2:
        tf += ev[0].elapsed_time(ev[1]); tb += ev[1].elapsed_time(ev[2]); to += ev[2].elapsed_time(ev[3])
print(f"batch {batch}: forward {tf / iters:.3f} ms  loss+backward {tb / iters:.3f} ms  adam {to / iters:.3f} ms")
