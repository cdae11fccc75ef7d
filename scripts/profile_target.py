"""Minimal workload for ncu: a few forwards of the hot path at one batch size (python scripts/profile_target.py B [reps])."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roko_b200.rnn_model import RNN, IN_SIZE, HIDDEN_SIZE, NUM_LAYERS  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
model = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS)
model.load_state_dict(torch.load(os.path.join(ROOT, "tests/golden/rand_seed1.pth"), map_location="cpu"))
model = model.to("cuda:0").eval()
g = torch.Generator(device="cuda:0").manual_seed(5)
x = torch.randint(0, 12, (B, 200, 90), dtype=torch.uint8, device="cuda:0", generator=g)
with torch.no_grad():
    for _ in range(reps):
        model.predict(x)
torch.cuda.synchronize()
print("done", B, reps)
