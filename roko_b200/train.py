"""Drop-in for the reference's ``roko/train.py`` entry point (same CLI, same ``.hdf5`` labelled
feature format, same ``.pth`` checkpoints), with ``model(x)`` and its backward running the B200
training kernels (roko_b200/csrc/train*.cu, gemm.cu, rec_bwd.cu) behind ``torch.autograd``.

    python -m roko_b200.train <train.hdf5|dir> <out_dir> [--val path] [--memory] [--t workers] [--b batch]

Reference behaviour mirrored (file:line in /root/reference/roko):
  * constants BATCH_SIZE 128, EPOCHS 100, LR 1e-4, PATIENCE 7                          train.py:12-15
  * datasets: every group except ``info`` / ``contigs``, ``examples`` + ``labels`` rows,
    a directory means all its ``*.hdf5``; lazy per-process handles or fully in memory   datasets.py:9-121
  * step: train mode, zero_grad, ``F.cross_entropy(model(x).transpose(1, 2), y)``,
    backward, Adam step                                                                 train.py:41-55
  * evaluation: eval mode, no_grad, accuracy over positions and mean loss               train.py:57-71
  * running average of the training loss (ignite RunningAverage, alpha 0.98)            train.py:69
  * early stopping on validation accuracy, patience 7; best-accuracy checkpoint
    ``rnn_model_<n>_acc=<score>.pth`` holding the state_dict, one file kept             train.py:73-84

ignite, tqdm and torchvision are not needed.  ``h5py`` is needed to read real files (absent in this
image; tests inject an in-memory stand-in).  Under ``torch.distributed`` (one process per GPU) ``--b`` is the
batch PER GPU (BASELINE.json config 5: 128 windows per GPU): the loader yields global batches of
``b x world`` windows, rank r takes windows [r b, (r+1) b), every rank's loss is the sum over its positions
divided by the GLOBAL position count, and one flat all-reduce (``dist.sum_gradients``) adds the gradients --
exactly the gradient of the mean loss over the global batch, also for a ragged last batch or an empty
shard.  Dropout masks differ per rank (the rank is mixed into the seed); rank 0 evaluates and writes
checkpoints.
"""
import argparse
import os

import numpy as np
import torch
import torch.nn.functional as F
from torch.utils.data import DataLoader, Dataset

from .rnn_model import RNN, IN_SIZE, HIDDEN_SIZE, NUM_LAYERS

BATCH_SIZE = 128
EPOCHS = 100
LR = 1e-4
PATIENCE = 7
RUNNING_ALPHA = 0.98          # ignite.metrics.RunningAverage default


def _h5(path=None):
    if isinstance(path, str) and path.startswith("synthetic://"):      # roko_b200/synth.py: a seeded stand-in with the same schema
        from . import synth
        return synth
    try:
        import h5py
        return h5py
    except ImportError as e:      # pragma: no cover - depends on the image
        raise RuntimeError("reading .hdf5 feature files needs h5py, which is not installed here") from e


def get_filenames(path):
    if path.startswith("synthetic://"):
        return [path]
    if os.path.isdir(path):
        return sorted(os.path.join(path, f) for f in os.listdir(path) if f.endswith(".hdf5"))
    return [path]


def _data_groups(fd):
    return [g for g in fd.keys() if g not in ("info", "contigs")]


class TrainToTensor:
    def __call__(self, sample):
        x, y = sample
        return torch.from_numpy(np.ascontiguousarray(x)), torch.from_numpy(np.ascontiguousarray(y))


class TrainDataset(Dataset):
    """Labelled windows read row by row from the files (lazy handles, one set per worker process)."""

    def __init__(self, path, transform=None, h5=None):
        self.filenames, self.transform, self._h5mod = get_filenames(path), transform, h5
        self.idx, self.fds = [], None
        for i, name in enumerate(self.filenames):
            fd = self._open(name)
            try:
                for g in _data_groups(fd):
                    self.idx.extend((i, g, j) for j in range(int(fd[g].attrs["size"])))
            finally:
                fd.close()

    def _open(self, name):
        return (self._h5mod or _h5(name)).File(name, "r")

    def __getitem__(self, i):
        if self.fds is None:
            self.fds = [self._open(n) for n in self.filenames]
        f, g, p = self.idx[i]
        group = self.fds[f][g]
        sample = (group["examples"][p], group["labels"][p])
        return self.transform(sample) if self.transform else sample

    def __len__(self):
        return len(self.idx)


class InMemoryTrainDataset(Dataset):
    """All labelled windows of the files as two contiguous arrays."""

    def __init__(self, path, transform=None, h5=None):
        xs, ys = [], []
        for name in get_filenames(path):
            fd = (h5 or _h5(name)).File(name, "r")
            try:
                for g in _data_groups(fd):
                    xs.append(np.asarray(fd[g]["examples"][:], dtype=np.uint8))
                    ys.append(np.asarray(fd[g]["labels"][:]))
            finally:
                fd.close()
        self.X = np.concatenate(xs) if xs else np.zeros((0, 200, 90), np.uint8)
        self.Y = np.concatenate(ys) if ys else np.zeros((0, 90), np.int64)
        assert len(self.X) == len(self.Y)
        self.transform = transform

    def __getitem__(self, i):
        sample = (self.X[i], self.Y[i])
        return self.transform(sample) if self.transform else sample

    def __len__(self):
        return len(self.X)


class SlabLoader:
    """Batches of an ``InMemoryTrainDataset`` without the per-item path of ``DataLoader``: one gather per batch
    (``index_select`` into two reusable staging buffers, pinned when the target is a CUDA device) instead of 128
    ``__getitem__`` calls and a collate, then an asynchronous copy to ``device``.  Same semantics as
    ``DataLoader(ds, batch_size, shuffle)``: a fresh permutation per epoch from ``generator``, the ragged last
    batch kept.  A staging buffer is rewritten only after the copy that last read it has completed (one CUDA
    event per buffer), so the host may run ahead of the device by a step without corrupting a batch."""

    def __init__(self, ds, batch_size, shuffle=False, generator=None, device="cpu", shard=None):
        # shard = (rank, per_rank_batch): data-parallel training gathers ONLY this rank's windows [rank b, (rank + 1) b) of every
        # global batch (all ranks walk the same permutation) and yields (x, y, global batch size)
        self.shard = shard
        self.x = torch.from_numpy(np.ascontiguousarray(ds.X))
        self.y = torch.from_numpy(np.ascontiguousarray(ds.Y))
        self.batch_size, self.shuffle, self.generator = int(batch_size), shuffle, generator
        self.device = torch.device(device)
        self.n = self.x.shape[0]
        pin = self.device.type == "cuda"
        self.bufs = [(torch.empty((self.batch_size,) + tuple(self.x.shape[1:]), dtype=self.x.dtype, pin_memory=pin),
                      torch.empty((self.batch_size,) + tuple(self.y.shape[1:]), dtype=self.y.dtype, pin_memory=pin))
                     for _ in range(2)]
        self.copied = [None, None]

    def __len__(self):
        return (self.n + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        order = torch.randperm(self.n, generator=self.generator) if self.shuffle else torch.arange(self.n)
        for k, lo in enumerate(range(0, self.n, self.batch_size)):
            idx = order[lo:lo + self.batch_size]
            n_global = idx.numel()
            if self.shard is not None:
                rank, b = self.shard
                idx = idx[min(rank * b, n_global):min((rank + 1) * b, n_global)]
            slot = k & 1
            bx, by = self.bufs[slot]
            m = idx.numel()
            if self.copied[slot] is not None:
                self.copied[slot].synchronize()                 # the device has finished reading this buffer
            torch.index_select(self.x, 0, idx, out=bx[:m])
            torch.index_select(self.y, 0, idx, out=by[:m])
            if self.device.type != "cuda":
                yield (bx[:m].clone(), by[:m].clone()) + ((n_global,) if self.shard is not None else ())
                continue
            dx = bx[:m].to(self.device, non_blocking=True)
            dy = by[:m].to(self.device, non_blocking=True)
            self.copied[slot] = torch.cuda.Event()
            self.copied[slot].record(torch.cuda.current_stream(self.device))
            yield (dx, dy) + ((n_global,) if self.shard is not None else ())


class EarlyStopping:
    """ignite.handlers.EarlyStopping(patience, score_function, trainer) with min_delta 0."""

    def __init__(self, patience):
        self.patience, self.best, self.counter = patience, None, 0

    def step(self, score):
        """True when training should stop."""
        if self.best is None or score > self.best:
            self.best, self.counter = score, 0
            return False
        self.counter += 1
        return self.counter >= self.patience


class BestCheckpoint:
    """ignite.handlers.ModelCheckpoint(out, 'rnn', score_function, score_name='acc', n_saved=1,
    require_empty=False) for one object named 'model': keeps the best-scoring state_dict only."""

    def __init__(self, dirname, prefix="rnn", name="model", score_name="acc"):
        self.dirname, self.prefix, self.name, self.score_name = dirname, prefix, name, score_name
        self.calls, self.best, self.path = 0, None, None
        os.makedirs(dirname, exist_ok=True)

    def step(self, score, module):
        self.calls += 1
        if self.best is not None and score <= self.best:
            return None
        path = os.path.join(self.dirname, f"{self.prefix}_{self.name}_{self.calls}_{self.score_name}={score:.7}.pth")
        torch.save({k: v.detach().cpu() for k, v in module.state_dict().items()}, path)
        if self.path and os.path.exists(self.path):
            os.remove(self.path)
        self.best, self.path = score, path
        return path


def _dist_info():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def evaluate(model, loader, device):
    """(accuracy over positions, mean cross-entropy) of ``model`` on ``loader`` -- train.py:57-71."""
    model.eval()
    correct = total = 0
    loss_sum, n_windows = 0.0, 0
    with torch.no_grad():
        for x, y in loader:
            x = x.to(device, non_blocking=True)
            y = y.to(device, non_blocking=True).long()
            out = model(x).transpose(1, 2)
            loss_sum += F.cross_entropy(out, y).item() * x.shape[0]
            n_windows += x.shape[0]
            correct += (out.argmax(dim=1) == y).sum().item()
            total += y.numel()
    return (correct / total if total else 0.0), (loss_sum / n_windows if n_windows else 0.0)


def train(train_path, out, val_path=None, mem=False, workers=0, batch_size=BATCH_SIZE, *, epochs=EPOCHS, lr=LR,
          patience=PATIENCE, model=None, device=None, h5=None, log=print, seed=None):
    """The reference's ``train`` (train.py:18-112).  Returns a dict with the history and the best checkpoint.

    ``model`` / ``device`` / ``h5`` are for tests and embedding; by default the model is
    ``roko_b200.RNN(500, 128, 3)`` on the current CUDA device."""
    from . import dist as rdist
    rank, world = _dist_info()
    data_class = InMemoryTrainDataset if mem else TrainDataset
    train_ds = data_class(train_path, transform=TrainToTensor(), h5=h5)
    val_ds = data_class(val_path, transform=TrainToTensor(), h5=h5) if val_path else None
    if seed is not None:
        torch.manual_seed(seed)
    shuffle_seed = int(torch.randint(0, 2 ** 31, (1,)).item()) if seed is None else int(seed)
    if world > 1:                                                           # every rank walks the same batches
        import torch.distributed as dist
        box = [shuffle_seed]
        dist.broadcast_object_list(box, src=0)
        shuffle_seed = box[0]
    gen = torch.Generator()
    gen.manual_seed(shuffle_seed)
    if world > 1:                                                           # shuffles agree across ranks, dropout masks must not
        torch.manual_seed(shuffle_seed * 1000003 + rank)
    global_batch = batch_size * world
    pin = torch.cuda.is_available()
    if mem:                                                                 # batch-granular gathers, no per-item collate
        train_dl = val_dl = None                                            # built once the device is known
    else:
        train_dl = DataLoader(train_ds, global_batch, shuffle=True, num_workers=workers, pin_memory=pin, generator=gen)
        val_dl = DataLoader(val_ds, batch_size, num_workers=workers, pin_memory=pin) if val_ds is not None else None

    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    device = torch.device(device)
    if mem:
        train_dl = SlabLoader(train_ds, global_batch, shuffle=True, generator=gen, device=device,
                              shard=(rank, batch_size) if world > 1 else None)
        val_dl = SlabLoader(val_ds, batch_size, device=device) if val_ds is not None else None
    if model is None:
        model = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS).to(device)          # raises without a B200: no CPU path
    if world > 1:
        rdist.broadcast_weights(model, src=0)                               # same start and same shuffles everywhere
    # the reference's Adam (train.py:43); on a GPU the single-kernel (fused) implementation of the same update
    optim = torch.optim.Adam(model.parameters(), lr=lr, fused=device.type == "cuda")
    if rank != 0:
        log = lambda *a, **k: None                                          # one log, as the reference's single process writes
    stopper, saver = EarlyStopping(patience), BestCheckpoint(out) if rank == 0 else None
    log(f"Device: {device}  ranks: {world}  train windows: {len(train_ds)}"
        + (f"  val windows: {len(val_ds)}" if val_ds is not None else ""))

    history = {"train_loss": [], "val_acc": [], "val_loss": [], "checkpoint": None, "epochs": 0, "epoch_s": [], "epoch_windows": []}
    import time
    running = None                       # ignite's RunningAverage, kept on the device: no host sync per step
    for epoch in range(1, epochs + 1):
        t_epoch, seen = time.perf_counter(), 0
        for i, item in enumerate(train_dl, 1):
            x, y = item[0], item[1]
            if len(item) == 3:                                              # the loader already took this rank's windows
                n_global = item[2]
            else:
                n_global = x.shape[0]
                if world > 1:                                               # this rank's windows of the global batch
                    lo, hi = min(rank * batch_size, n_global), min((rank + 1) * batch_size, n_global)
                    x, y = x[lo:hi], y[lo:hi]
            seen += int(n_global)
            x = x.to(device, non_blocking=True)
            y = y.to(device, non_blocking=True).long()
            model.train()
            model.zero_grad()
            loss = None
            if x.shape[0]:
                # sum over my positions / GLOBAL position count: the all-reduced SUM is the gradient of the global mean
                loss = F.cross_entropy(model(x).transpose(1, 2), y, reduction="sum") / (n_global * y.shape[1])
                loss.backward()
            if world > 1:
                rdist.sum_gradients(model)
            optim.step()
            if loss is not None:
                v = loss.detach() * (n_global / x.shape[0])                 # this rank's mean loss, for the log line
                running = v.clone() if running is None else running.mul_(RUNNING_ALPHA).add_(v, alpha=1.0 - RUNNING_ALPHA)
            if i % 100 == 0:
                log(f"ITERATION {i}/{len(train_dl)} - loss: {float(running) if running is not None else None}")
        history["train_loss"].append(float(running) if running is not None else None)   # (the float() synchronises the device)
        history["epochs"] = epoch
        history["epoch_s"].append(time.perf_counter() - t_epoch)
        history["epoch_windows"].append(seen)
        log(f"Epoch {epoch}: {seen} windows in {history['epoch_s'][-1]:.2f} s = {seen / history['epoch_s'][-1]:,.0f} windows/s over {world} GPU(s)")
        if hasattr(model, "check_codes"):
            model.check_codes()              # nn.Embedding would have raised IndexError on a code outside 0..11
        if val_dl is None:
            continue
        stop = False
        if rank == 0:
            acc, vloss = evaluate(model, val_dl, device)
            history["val_acc"].append(acc)
            history["val_loss"].append(vloss)
            log(f"Val epoch: {epoch}, acc: {acc}, loss: {vloss}")
            path = saver.step(acc, model)
            if path:
                history["checkpoint"] = path
            stop = stopper.step(acc)
        if world > 1:
            import torch.distributed as dist
            flag = torch.tensor([1 if stop else 0], device=device if device.type == "cuda" else "cpu")
            dist.broadcast(flag, src=0)
            stop = bool(flag.item())
        if stop:
            log(f"EarlyStopping: no improvement of val_acc in {patience} evaluations")
            break
    return history


def main(argv=None):
    parser = argparse.ArgumentParser(description="train the roko consensus network on B200 (reference roko/train.py)")
    parser.add_argument("train", type=str)
    parser.add_argument("out", type=str)
    parser.add_argument("--val", type=str, default=None)
    parser.add_argument("--memory", action="store_true", default=False)
    parser.add_argument("--t", type=int, default=0)
    parser.add_argument("--b", type=int, default=BATCH_SIZE)
    parser.add_argument("--epochs", type=int, default=EPOCHS, help="(addition) stop after this many epochs; the reference runs 100 with early stopping")
    args = parser.parse_args(argv)
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:                # launched by torchrun: one process per GPU
        import torch.distributed as dist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    hist = train(args.train, args.out, args.val, args.memory, args.t, args.b, epochs=args.epochs)
    if "RANK" in os.environ and "WORLD_SIZE" in os.environ:
        import torch.distributed as dist
        dist.destroy_process_group()
    return hist


if __name__ == "__main__":
    main()
