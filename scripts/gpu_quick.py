"""Quick on-GPU look: stage timings through the C ABI at a few batch sizes + FP32 peak."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from roko_b200 import _cabi  # noqa: E402
from roko_b200.rnn_model import RNN, IN_SIZE, HIDDEN_SIZE, NUM_LAYERS  # noqa: E402
from roko_b200.synth import uniform_windows  # noqa: E402

STAGES = ["front", "proj0", "rec0", "proj1", "rec1", "proj2", "rec2", "head"]


def stage_times(model, x, iters=10):
    h = model._handle(x.device)
    n = x.shape[0]
    ws = torch.empty(h.lib.roko_b200_workspace_bytes(n), dtype=torch.uint8, device=x.device)
    labels = torch.empty((n, 90), dtype=torch.uint8, device=x.device)
    ms = (ctypes.c_float * 8)()
    stream = torch.cuda.current_stream().cuda_stream
    for it in (2, iters):
        _cabi.check(h.lib.roko_b200_forward_timed(h.ptr, x.data_ptr(), n, labels.data_ptr(), ws.data_ptr(),
                                                   ws.numel(), stream, it, ms))
    return list(ms)


def main():
    lib = _cabi.lib()
    tf = ctypes.c_double()
    _cabi.check(lib.roko_b200_measure_fp32_peak(0, ctypes.byref(tf)))
    print(f"fp32 FFMA peak measured: {tf.value:.1f} TFLOP/s")
    sd = torch.load(os.path.join(ROOT, "tests/golden/rand_seed1.pth"), map_location="cpu")
    model = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS)
    model.load_state_dict(sd)
    model = model.to("cuda:0").eval()
    for n in [int(a) for a in sys.argv[1:]] or [128, 296, 1024]:
        x = torch.from_numpy(uniform_windows(n, seed=3)).cuda()
        ms = stage_times(model, x)
        tot = sum(ms)
        print(f"B={n}: total {tot:.3f} ms -> {n / tot * 1e3:,.0f} windows/s | " +
              " ".join(f"{s}={v:.3f}" for s, v in zip(STAGES, ms)))
        # end-to-end forward via the public call, 20 reps on one stream
        with torch.no_grad():
            for _ in range(3):
                model.predict(x)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
            e0.record()
            for _ in range(20):
                model.predict(x)
            e1.record()
            torch.cuda.synchronize()
            t = e0.elapsed_time(e1) / 20
        print(f"      predict(): {t:.3f} ms/batch -> {n / t * 1e3:,.0f} windows/s")


if __name__ == "__main__":
    main()
