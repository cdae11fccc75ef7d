"""GPU, two devices: the drop-in RNN under torch's own DistributedDataParallel (the reference's class is
wrapped the same way by users; SURVEY.md 8b "wrappable by nn.DataParallel/DDP") and under
roko_b200.dist.average_gradients -- both must leave every rank with the mean of the per-rank gradients,
i.e. the gradient of the concatenated batch.  Skipped on a single-GPU box."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    import torch.nn.functional as F
    from torch.nn.parallel import DistributedDataParallel as DDP
    from roko_b200 import dist as rdist
    from roko_b200.rnn_model import RNN
    from roko_b200.synth import structured_windows

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        sd = torch.load(os.path.join(root, "tests", "golden", "rand_seed1.pth"), map_location="cpu")
        x, y = structured_windows(4, seed=7217, return_truth=True)
        xs = torch.from_numpy(x[2 * rank:2 * rank + 2]).cuda()
        ys = torch.from_numpy(y[2 * rank:2 * rank + 2].astype(np.int64)).cuda()

        def grads(wrap):
            m = RNN(500, 128, 3)
            m.load_state_dict(sd)
            m = m.cuda().eval()                      # eval + autograd: no dropout, deterministic
            net = DDP(m, device_ids=[rank]) if wrap else m
            F.cross_entropy(net(xs).transpose(1, 2), ys).backward()
            if not wrap:
                rdist.average_gradients(m)
            return torch.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu().numpy()

        g_ddp, g_avg = grads(True), grads(False)
        q.put((rank, g_ddp, g_avg))
    finally:
        dist.destroy_process_group()


def test_ddp_and_average_gradients_match_full_batch(seed1_state):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    import torch.nn.functional as F
    from roko_b200.rnn_model import RNN
    from roko_b200.synth import structured_windows
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=240) for _ in procs), key=lambda r: r[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single process, whole batch of 4: the mean over ranks of half-batch mean losses is the full-batch mean loss
    x, y = structured_windows(4, seed=7217, return_truth=True)
    m = RNN(500, 128, 3)
    m.load_state_dict(seed1_state)
    m = m.to("cuda:0").eval()
    F.cross_entropy(m(torch.from_numpy(x).cuda()).transpose(1, 2), torch.from_numpy(y.astype(np.int64)).cuda()).backward()
    full = torch.cat([p.grad.reshape(-1) for p in m.parameters()]).cpu().numpy()
    scale = np.abs(full).max()
    for rank, g_ddp, g_avg in res:
        assert np.abs(g_ddp - full).max() <= 1e-4 * scale, rank
        assert np.abs(g_avg - full).max() <= 1e-4 * scale, rank
    assert np.array_equal(res[0][1], res[1][1])          # DDP leaves identical gradients on both ranks


def _infer_worker(rank, world, port, q, n_total):
    """One rank of sharded inference: contiguous window range, NCCL weight broadcast, label gather to rank 0."""
    import torch.distributed as dist
    from roko_b200 import dist as rdist
    from roko_b200.rnn_model import RNN
    from roko_b200.synth import structured_windows
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
        m = RNN(500, 128, 3)
        if rank == 0:                                   # only rank 0 has the weights; the others get them over NCCL
            m.load_state_dict(torch.load(os.path.join(root, "tests", "golden", "rand_seed1.pth"), map_location="cpu"))
        m = m.cuda().eval().requires_grad_(False)
        rdist.broadcast_weights(m, src=0)
        x = structured_windows(n_total, seed=4242)
        lo, hi = rdist.shard_range(n_total, rank, world)
        local = m.predict(torch.from_numpy(x[lo:hi]).cuda())
        got = rdist.gather_labels(local, n_total)
        if rank == 0:
            q.put(got.cpu().numpy())
    finally:
        dist.destroy_process_group()


def test_sharded_inference_equals_single_gpu(cuda_model):
    """shard -> gather equals the single-GPU output byte for byte (ragged shards: 301 windows)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from roko_b200.synth import structured_windows
    n_total = 301
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_infer_worker, args=(r, 2, port, q, n_total)) for r in range(2)]
    for p in procs:
        p.start()
    got = q.get(timeout=240)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    single = cuda_model.predict(torch.from_numpy(structured_windows(n_total, seed=4242)).cuda()).cpu().numpy()
    assert got.shape == single.shape and np.array_equal(got, single)
