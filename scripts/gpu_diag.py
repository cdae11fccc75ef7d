"""On-GPU kernel A/B diagnostic: every kernel configuration against the fp32-exact FFMA configuration, tap by tap,
plus stage timings and per-call host overhead.  Each configuration runs in its own subprocess under a timeout, so a
kernel that deadlocks costs one line of output, not the GPU call.

    python scripts/gpu_diag.py            # all configurations
    python scripts/gpu_diag.py one '{"proj": 4, "rec": 2}'      # one configuration, in-process
"""
import ctypes
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
STAGES = ["front", "proj0", "rec0", "proj1", "rec1", "proj2", "rec2", "head"]

CONFIGS = [
    {"name": "ffma/ffma (reference)", "proj": 0, "rec_tc_min": 0, "front": 0},
    {"name": "proj fp16 + rec ffma", "proj": 4, "rec_tc_min": 0, "front": 0},
    {"name": "proj ffma + rec fp16", "proj": 0, "rec": 2, "rec_tc_min": 32, "front": 0},
    {"name": "proj fp16 + rec fp16 + front mma.sync", "proj": 4, "rec": 2, "rec_tc_min": 32, "front": 0},
    {"name": "proj tf32 + rec tf32 + front mma.sync (round 1)", "proj": 3, "rec": 1, "rec_tc_min": 64, "front": 0},
    {"name": "default: proj fp16 + rec fp16 + front tcgen05", "proj": 4, "rec": 2, "rec_tc_min": 32, "front": 1},
]


def one(cfg):
    import numpy as np
    import torch
    from roko_b200 import _cabi
    from roko_b200.rnn_model import RNN, IN_SIZE, HIDDEN_SIZE, NUM_LAYERS
    from roko_b200.synth import structured_windows, uniform_windows

    def make(opts):
        m = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS)
        m.load_state_dict(torch.load(os.path.join(ROOT, "tests/golden/rand_seed1.pth"), map_location="cpu"))
        m = m.to("cuda:0").eval().requires_grad_(False)
        for k, v in opts.items():
            if k != "name":
                m.set_option(k, v)
        return m

    ref = make(CONFIGS[0])
    m = make(cfg)
    out = {"name": cfg["name"]}
    for n in (37, 128):
        x = torch.from_numpy(structured_windows(n, seed=900 + n)).cuda()
        a, b = ref.forward_taps(x), m.forward_taps(x)
        out[f"taps@{n}"] = {k: float((a[k].float() - b[k].float()).abs().max()) for k in ("front", "gru_l0", "gru_l1", "gru_l2", "logits")}
        out[f"labels_equal@{n}"] = bool(torch.equal(a["labels"], b["labels"]))
    try:
        m.check_codes()
    except Exception as ex:
        out["check_codes"] = repr(ex)[:120]
    h = m._handle(torch.device("cuda:0"))
    for n in [int(v) for v in os.environ.get("DIAG_SIZES", "128,2368").split(",")]:
        x = torch.from_numpy(uniform_windows(n, seed=3)).cuda()
        ws = torch.empty(h.lib.roko_b200_workspace_bytes(n), dtype=torch.uint8, device="cuda:0")
        labels = torch.empty((n, 90), dtype=torch.uint8, device="cuda:0")
        ms = (ctypes.c_float * 8)()
        for it in (2, 10):
            _cabi.check(h.lib.roko_b200_forward_timed(h.ptr, x.data_ptr(), n, labels.data_ptr(), ws.data_ptr(), ws.numel(),
                                                       torch.cuda.current_stream().cuda_stream, it, ms))
        out[f"stage_ms@{n}"] = {s: round(float(v), 4) for s, v in zip(STAGES, ms)}
        out[f"windows_per_s@{n}"] = round(n / sum(ms) * 1e3)
    # per-call host cost of predict(): tiny batch so the GPU never backs up, graphs on and off
    x1 = torch.from_numpy(uniform_windows(2, seed=4)).cuda()
    y1 = torch.empty((2, 90), dtype=torch.uint8, device="cuda:0")
    s = torch.cuda.Stream()
    for graphs in (1, 0):
        m.set_option("graphs", graphs)
        with torch.cuda.stream(s), torch.no_grad():
            for _ in range(20):
                m.predict(x1, out=y1)
            s.synchronize()
            t0 = time.perf_counter()
            for _ in range(300):
                m.predict(x1, out=y1)
            host = (time.perf_counter() - t0) / 300
            s.synchronize()
            wall = (time.perf_counter() - t0) / 300
        out[f"predict_host_us(graphs={graphs})"] = round(host * 1e6, 1)
        out[f"predict_wall_us(graphs={graphs})"] = round(wall * 1e6, 1)
    print("DIAG " + json.dumps(out), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "one":
        one(json.loads(sys.argv[2]))
        return
    only = sys.argv[1:] or None
    for cfg in CONFIGS:
        if only and not any(o in cfg["name"] for o in only):
            continue
        t0 = time.time()
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "one", json.dumps(cfg)], capture_output=True, text=True, timeout=150)
            lines = [l for l in p.stdout.splitlines() if l.startswith("DIAG ")]
            print(lines[-1] if lines else f"DIAG-FAIL {cfg['name']} rc={p.returncode} :: {p.stderr.strip()[-600:]}", flush=True)
        except subprocess.TimeoutExpired:
            print(f"DIAG-TIMEOUT {cfg['name']} after {time.time() - t0:.0f} s (kernel hang?)", flush=True)


if __name__ == "__main__":
    main()
