"""CPU: the N>1 host logic (sharding, weight broadcast, label gather) under gloo, world_size 2."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from roko_b200.dist import shard_range, shard_sizes


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 1000, 1_000_000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = shard_sizes(n, world)
            assert sum(sizes) == n and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import roko_b200.dist as rd
        from roko_b200.rnn_model import RNN
        from oracle import roko_oracle as O
        from roko_b200.synth import structured_windows

        torch.manual_seed(100 + rank)                     # ranks start with DIFFERENT weights
        model = RNN(500, 128, 3)
        nbytes = rd.broadcast_weights(model, src=0)
        flat = torch.cat([p.detach().reshape(-1) for _, p in sorted(model.named_parameters())])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(flat, ref))

        # every rank labels its contiguous shard (the oracle stands in for the CUDA path on CPU)
        x = structured_windows(n_total, seed=42)
        lo, hi = rd.shard_range(n_total, rank, world)
        w = {k: v.detach().numpy() for k, v in model.state_dict().items()}
        local = torch.from_numpy(O.predict(x[lo:hi], w))
        out = rd.gather_labels(local, n_total)
        if rank == 0:
            full = O.predict(x, w)
            q.put((same, nbytes, bool(np.array_equal(out.numpy(), full)), tuple(out.shape)))
        else:
            assert out is None
            q.put((same, nbytes, True, None))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_total", [5, 6])               # ragged and even shards
def test_broadcast_and_gather_world2(n_total):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=240) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for same, nbytes, ok, shape in res:
        assert same and ok and nbytes == 1099731 * 4
    assert any(shape == (n_total, 90) for _, _, _, shape in res)


def _train_worker(rank, world, port, outdir, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import roko_b200.dist as rd
        from roko_b200 import train as T
        from roko_b200.synth import structured_windows
        from tests import fake_h5
        from tests.test_train_host import TinyModel

        # (1) average_gradients: mean over ranks, parameters without a local gradient count as zeros
        torch.manual_seed(7)
        m = TinyModel()
        m.fc.weight.grad = torch.full_like(m.fc.weight, float(rank + 1))
        if rank == 0:
            m.fc.bias.grad = torch.full_like(m.fc.bias, 4.0)
        nbytes = rd.average_gradients(m)
        ok_avg = bool(torch.allclose(m.fc.weight.grad, torch.full_like(m.fc.weight, 1.5))
                      and torch.allclose(m.fc.bias.grad, torch.full_like(m.fc.bias, 2.0)))

        # (2) two ranks with 4 windows per rank and step == one process with 8 windows per step, including the ragged
        # last batch of 30 windows (6 = 4 + 2: shards weigh in by their size, not as a mean of means)
        x, y = structured_windows(30, seed=21, return_truth=True)
        pos = np.zeros((30, 90, 2), np.int64)
        fake_h5.register("mem://dist_train", {"c": "ACGT"}, [("c_0", "c", pos, x, y)])
        torch.manual_seed(100 + rank)                     # different initial weights: rank 0's must win
        model = TinyModel()
        hist = T.train("mem://dist_train", outdir, "mem://dist_train", mem=True, batch_size=4, epochs=2, lr=1e-2,
                       model=model, device="cpu", h5=fake_h5, log=lambda *_: None, seed=5)
        flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
        q.put((rank, ok_avg, nbytes, flat.numpy(), hist["checkpoint"]))
    finally:
        dist.destroy_process_group()


def test_data_parallel_training_world2(tmp_path):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_train_worker, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert all(r[1] for r in res) and res[0][2] == (5 * 12 + 5) * 4
    assert np.allclose(res[0][3], res[1][3], atol=1e-7)          # ranks stay in lock step
    assert res[0][4] is not None and res[1][4] is None            # rank 0 alone writes checkpoints

    # single-process run from rank 0's start with the global batch (8): size-weighted shard gradients == whole-batch gradient
    from roko_b200 import train as T
    from roko_b200.synth import structured_windows
    from tests import fake_h5
    from tests.test_train_host import TinyModel
    x, y = structured_windows(30, seed=21, return_truth=True)
    fake_h5.register("mem://dist_train", {"c": "ACGT"}, [("c_0", "c", np.zeros((30, 90, 2), np.int64), x, y)])
    torch.manual_seed(100)
    model = TinyModel()
    T.train("mem://dist_train", str(tmp_path / "single"), None, mem=True, batch_size=8, epochs=2, lr=1e-2,
            model=model, device="cpu", h5=fake_h5, log=lambda *_: None, seed=5)
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()]).numpy()
    assert np.allclose(flat, res[0][3], atol=2e-5)
