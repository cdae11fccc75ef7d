"""BASELINE.json configs[3]: N synthetic windows sharded across the GPUs of one box through the PRODUCT entry point
(roko_b200.inference.infer_fast under torchrun): NCCL weight broadcast, contiguous window shards, label gather to rank 0,
vote + stitch + FASTA on rank 0.  Prints one JSON line on rank 0.

    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 scripts/run_config4.py 1000000 [out.json]
    python scripts/run_config4.py 100000            # single GPU
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from roko_b200 import inference, synth  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100000
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    out = os.path.join(tempfile.gettempdir(), f"config4_{os.getpid()}.fasta")
    path = f"synthetic://{n}?cache=1"
    pth = os.path.join(ROOT, "tests", "golden", "rand_seed1.pth")
    t0 = time.perf_counter()
    inference.infer_fast(path, pth, out, workers=0, batch_size=1024, h5=synth, chunk=16384)       # cold: generates this rank's windows
    cold = time.perf_counter() - t0
    if world > 1:
        torch.distributed.barrier()
    t0 = time.perf_counter()
    stats = {}
    recs = inference.infer_fast(path, pth, out, workers=0, batch_size=1024, h5=synth, chunk=16384, stats=stats)  # warm: windows come from memory
    dt = time.perf_counter() - t0
    if rank == 0:
        line = {"config": "BASELINE configs[3]: synthetic windows sharded over the GPUs of one box through roko_b200.inference.infer_fast",
                "windows": n, "n_gpus": world, "wall_s": dt, "windows_per_s_end_to_end": n / dt, "cold_wall_s": cold,
                "phases_s_rank0": {k: round(v, 3) for k, v in stats.items()}, "contigs": len(recs), "consensus_bases": sum(len(s) for _, s in recs),
                "note": "wall_s is the second pass (windows served from host memory, like a page-cached .hdf5): model construction, NCCL weight "
                        "broadcast, slab reads, pinned staging, H2D, the model path, D2H, the NCCL label gather and rank 0's vote / stitch / FASTA "
                        "write; cold_wall_s additionally generates the synthetic windows (numpy PCG64, ~0.26 GB/s per rank)"}
        print(json.dumps(line), flush=True)
        if len(sys.argv) > 2:
            with open(sys.argv[2], "w") as f:
                json.dump(line, f)
        os.remove(out)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
