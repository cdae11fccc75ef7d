"""Drop-in for the reference's ``roko/inference.py`` entry point (same CLI, same ``.hdf5`` feature
and ``.pth`` weight formats, same FASTA output), with the model call replaced by the B200 path.

    python -m roko_b200.inference <data.hdf5> <model.pth> <out.fasta> [--t workers] [--b batch]
    torchrun --nproc-per-node 8 -m roko_b200.inference <data.hdf5> <model.pth> <out.fasta>      # windows sharded over 8 GPUs

Reference behaviour mirrored (file:line in /root/reference/roko/inference.py):
  * alphabet / decoding                                              :14-17
  * ``InferenceDataset``: flat index over every group except ``contigs`` (``attrs['size']``),
    contig drafts from ``/contigs/<name>.attrs['seq'|'len']``, lazy per-process file handle   :27-87
  * ``infer``: load state_dict, eval, batches in file order, argmax labels, one vote per
    (contig, (rpos, ins)) and label                                   :90-127
  * stitching: sort positions, drop leading insertion slots, majority base per position with
    ``Counter.most_common`` tie behaviour (first label seen wins), skip '*', splice the draft's
    prefix and suffix                                                 :129-151
  * FASTA via ``SeqIO.write`` of ``SeqRecord(seq, id=contig)``        :149-154

Two drivers share the dataset / vote / stitch code: ``infer`` keeps the reference's per-batch loop
(DataLoader of single windows, one model call per batch); ``infer_fast`` (the CLI default) reads
``examples[i:i+n]`` slabs per group, pushes them through ``RNN.predict_host`` (pinned staging, batches
coalesced into device passes) and votes with dense scatter-adds on the GPU.  Both produce the same FASTA.

Differences, all on the host side of the boundary: windows stay uint8 end to end (the reference
widens to int64 before the copy, :113); labels come back as uint8 from the fused argmax; the
per-position Python ``Counter`` loop (:119-124, ~44 ms per 128-window batch) is a vectorised
numpy scatter with the same tie rule.  ``h5py`` is needed to read real files (it is not in this
image; tests inject an in-memory stand-in); biopython is not needed.
"""
import argparse

import numpy as np
import torch
from torch.utils.data import DataLoader, Dataset

from .rnn_model import RNN, IN_SIZE, HIDDEN_SIZE, NUM_LAYERS

GAP = "*"
ALPHABET = "ACGT" + GAP + "N"
encoding = {v: i for i, v in enumerate(ALPHABET)}
decoding = {v: k for k, v in encoding.items()}
N_LABELS = 5          # classes the network emits: A C G T *
MAX_INS = 3           # insertion slots per reference position (reference include/generate.h:22)


def _h5(path=None):
    if isinstance(path, str) and path.startswith("synthetic://"):      # roko_b200/synth.py: a seeded stand-in with the same schema
        from . import synth
        return synth
    try:
        import h5py
        return h5py
    except ImportError as e:      # pragma: no cover - depends on the image
        raise RuntimeError("reading .hdf5 feature files needs h5py, which is not installed here") from e


class ToTensor:
    def __call__(self, sample):
        contig, position, x = sample
        return contig, position, torch.from_numpy(np.ascontiguousarray(x))


class InferenceDataset(Dataset):
    """Flat view of the windows of a feature file (schema: SURVEY.md App. C, reference data.py:40-48)."""

    def __init__(self, path, transform=None, h5=None):
        self.path, self.transform, self._h5mod = path, transform, h5
        self.contigs, self.idx, self.size, self.fd = {}, [], 0, None
        fd = self._open()
        try:
            for g in fd.keys():
                if g == "contigs":
                    continue
                n = int(fd[g].attrs["size"])
                self.idx.extend((g, j) for j in range(n))
                self.size += n
            for k in fd["contigs"]:
                grp = fd["contigs"][k]
                self.contigs[str(k)] = (grp.attrs["seq"], grp.attrs["len"])
        finally:
            fd.close()

    def _open(self):
        return (self._h5mod or _h5(self.path)).File(self.path, "r")

    def __getitem__(self, i):
        if self.fd is None:
            self.fd = self._open()          # lazily, once per DataLoader worker process
        g, p = self.idx[i]
        group = self.fd[g]
        sample = (group.attrs["contig"], group["positions"][p], group["examples"][p])
        return self.transform(sample) if self.transform else sample

    def __len__(self):
        return self.size

    def close_fd(self):
        if self.fd is not None:
            self.fd.close()
            self.fd = None


class VoteTable:
    """Votes per (contig, (rpos, ins)) with ``Counter`` semantics, vectorised.

    ``counts[slot, label]`` and the global sequence number of each label's first vote are kept so
    that ties resolve exactly like ``Counter.most_common(1)`` in the reference (:141): among the
    labels with the highest count the one that was voted first wins.
    """

    def __init__(self):
        self.tables = {}        # contig -> (slot_keys dict grows lazily) ; implemented as sparse dict of arrays
        self.seq = 0

    def add(self, contig, pos, labels):
        """pos (n,2) int64 [(rpos, ins)], labels (n,) uint8 -- in window/position order."""
        t = self.tables.get(contig)
        if t is None:
            t = self.tables[contig] = {"keys": np.empty(0, np.int64), "counts": np.zeros((0, N_LABELS), np.int32),
                                       "first": np.zeros((0, N_LABELS), np.int64)}
        key = pos[:, 0].astype(np.int64) * (MAX_INS + 1) + pos[:, 1].astype(np.int64)
        new = np.setdiff1d(key, t["keys"])                       # sorted unique
        if new.size:
            keys = np.concatenate([t["keys"], new])
            order = np.argsort(keys, kind="stable")
            t["keys"] = keys[order]
            t["counts"] = np.concatenate([t["counts"], np.zeros((new.size, N_LABELS), np.int32)])[order]
            t["first"] = np.concatenate([t["first"], np.full((new.size, N_LABELS), np.iinfo(np.int64).max)])[order]
        slot = np.searchsorted(t["keys"], key)
        lab = labels.astype(np.int64)
        np.add.at(t["counts"], (slot, lab), 1)
        np.minimum.at(t["first"], (slot, lab), self.seq + np.arange(key.size, dtype=np.int64))
        self.seq += key.size

    def consensus(self, contig):
        """[(rpos, ins)] sorted and the winning label per position."""
        t = self.tables[contig]
        counts, first = t["counts"], t["first"]
        best = counts.max(axis=1, keepdims=True)
        cand = np.where(counts == best, first, np.iinfo(np.int64).max)
        winner = cand.argmin(axis=1)
        keys = t["keys"]
        return np.stack([keys // (MAX_INS + 1), keys % (MAX_INS + 1)], axis=1), winner


_DECODE_LUT = np.frombuffer(ALPHABET.encode(), dtype="S1")


def stitch(contig_seq, positions, winners):
    """inference.py:129-147 for one contig: positions sorted [(rpos, ins)], winners label ids.
    Vectorised (a byte look-up and a mask instead of a Python loop per base: a 5 Mbp contig stitches in milliseconds)."""
    keep = np.flatnonzero(positions[:, 1] == 0)
    if keep.size == 0:
        raise IndexError("no reference-anchored position for contig")     # the reference raises here too (pos_sorted[0])
    start = keep[0]                                                       # dropwhile(ins != 0)
    positions, winners = positions[start:], np.asarray(winners[start:], dtype=np.int64)
    first, last = int(positions[0, 0]), int(positions[-1, 0])
    body = _DECODE_LUT[winners[winners != encoding[GAP]]].tobytes().decode("ascii")
    return contig_seq[:first] + body + contig_seq[last + 1:]


def write_fasta(records, path):
    """What ``SeqIO.write(records, f, 'fasta')`` emits for ``SeqRecord(Seq(s), id=c)``: the default
    description ``<unknown description>`` follows the id, sequence wrapped at 60 columns."""
    with open(path, "w") as f:
        for name, seq in records:
            f.write(f">{name} <unknown description>\n")
            for i in range(0, len(seq), 60):
                f.write(seq[i:i + 60] + "\n")


def infer(data, model_path, out, workers=0, batch_size=128, h5=None, device=None):
    if not torch.cuda.is_available():
        raise RuntimeError("roko_b200.inference needs a CUDA (sm_100a) device: the model path has no CPU fallback")
    device = torch.device(device or "cuda:0")

    model = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS).to(device)
    model.load_state_dict(torch.load(model_path, map_location=device))
    model.eval()

    dataset = InferenceDataset(data, transform=ToTensor(), h5=h5)
    dataloader = DataLoader(dataset, batch_size=batch_size, shuffle=False, num_workers=workers)
    votes = VoteTable()

    print("Inference started")
    with torch.no_grad():
        for i, (c, pos, x) in enumerate(dataloader):
            y = model.predict(x.to(device, non_blocking=True)).cpu().numpy()      # uint8 labels, fused argmax
            pos = pos.numpy()
            c = np.asarray(c)
            for contig in dict.fromkeys(c.tolist()):                              # contigs in order of appearance
                sel = np.flatnonzero(c == contig)
                votes.add(contig, pos[sel].reshape(-1, 2), y[sel].reshape(-1))
            if (i + 1) % 100 == 0:
                print(f"{i + 1} batches processed")

    model.check_codes()                       # nn.Embedding would have raised IndexError on a code outside 0..11
    records = []
    for contig in votes.tables:
        positions, winners = votes.consensus(contig)
        records.append((contig, stitch(dataset.contigs[contig][0], positions, winners)))
    write_fasta(records, out)
    return records


class DenseVoteTable:
    """Same ``Counter`` semantics as ``VoteTable`` but dense and in torch (CPU or CUDA tensors): per contig
    ``counts[(rpos - base) * 4 + ins, label]`` by scatter-add and the sequence number of each label's first vote by
    scatter-amin.

    A table covers only the position RANGE its contig's windows actually touch (it grows, with slack, when a
    later group reaches outside), costs 60 B per slot (240 B per draft base in range), and is dropped by
    ``release`` once the contig is stitched -- so the footprint follows the contig being processed, not the
    genome.  A contig whose range would exceed ``budget_bytes`` is handed to the sparse ``VoteTable`` instead
    (the reference's Counter path has no such limit either: roko/inference.py:101)."""

    SLOT_BYTES = N_LABELS * (4 + 8)

    def __init__(self, device, budget_bytes=4 << 30):
        self.device = torch.device(device)
        self.tables = {}
        self.sparse = VoteTable()
        self.seq = 0
        self.budget = int(budget_bytes)

    def _alloc(self, base, span):
        slots = span * (MAX_INS + 1)
        return {"base": base, "span": span,
                "counts": torch.zeros(slots * N_LABELS, dtype=torch.int32, device=self.device),
                "first": torch.full((slots * N_LABELS,), torch.iinfo(torch.int64).max, dtype=torch.int64, device=self.device)}

    def _cover(self, contig, lo, hi):
        """Make contig's table cover reference positions [lo, hi]; returns it, or None once the contig is sparse."""
        t = self.tables.get(contig)
        if t is None and contig in self.sparse.tables:
            return None
        if t is not None and t["base"] <= lo and hi < t["base"] + t["span"]:
            return t
        nlo = lo if t is None else min(lo, t["base"])
        nhi = hi if t is None else max(hi, t["base"] + t["span"] - 1)
        slack = 0 if t is None else max(1 << 16, (nhi - nlo + 1) // 2)          # amortise growth in file order
        nlo2, nhi2 = (nlo, nhi + slack) if t is None or lo >= t["base"] else (max(0, nlo - slack), nhi)
        span = nhi2 - nlo2 + 1
        if span * (MAX_INS + 1) * self.SLOT_BYTES > self.budget:
            self._spill(contig)
            return None
        n = self._alloc(nlo2, span)
        if t is not None:
            off = (t["base"] - nlo2) * (MAX_INS + 1) * N_LABELS
            n["counts"][off:off + t["counts"].numel()] = t["counts"]
            n["first"][off:off + t["first"].numel()] = t["first"]
        self.tables[contig] = n
        return n

    def _spill(self, contig):
        """Move a contig to the sparse table (keeps counts and first-vote order)."""
        t = self.tables.pop(contig, None)
        st = self.sparse.tables.setdefault(contig, {"keys": np.empty(0, np.int64), "counts": np.zeros((0, N_LABELS), np.int32),
                                                    "first": np.zeros((0, N_LABELS), np.int64)})
        if t is not None:
            counts = t["counts"].view(-1, N_LABELS)
            voted = torch.nonzero(counts.sum(dim=1) > 0)[:, 0]
            st["keys"] = (voted + t["base"] * (MAX_INS + 1)).cpu().numpy()
            st["counts"] = counts[voted].cpu().numpy()
            st["first"] = t["first"].view(-1, N_LABELS)[voted].cpu().numpy()

    def add(self, contig, contig_len, pos, labels):
        """pos (n,2) int64 tensor [(rpos, ins)], labels (n,) uint8 tensor, in window/position order.
        ``contig_len`` is informative only: tables are sized from the positions themselves."""
        n = labels.numel()
        if n == 0:
            return
        pos = torch.as_tensor(pos)
        if int(pos[:, 1].max()) > MAX_INS or int(pos.min()) < 0:
            raise IndexError("position outside (rpos >= 0, 0 <= ins <= %d)" % MAX_INS)
        lo, hi = int(pos[:, 0].min()), int(pos[:, 0].max())
        t = self._cover(contig, lo, hi)
        if t is None:                                                     # sparse contig: numpy path, shared sequence numbers
            self.sparse.seq = self.seq
            self.sparse.add(contig, pos.cpu().numpy(), labels.cpu().numpy())
            self.seq += n
            return
        pos = pos.to(self.device, torch.int64)
        idx = ((pos[:, 0] - t["base"]) * (MAX_INS + 1) + pos[:, 1]) * N_LABELS + labels.to(self.device, torch.int64)
        t["counts"].scatter_add_(0, idx, torch.ones_like(idx, dtype=torch.int32))
        order = torch.arange(self.seq, self.seq + n, dtype=torch.int64, device=self.device)
        t["first"].scatter_reduce_(0, idx, order, reduce="amin", include_self=True)
        self.seq += n

    def contigs(self):
        return list(dict.fromkeys(list(self.tables) + list(self.sparse.tables)))

    def consensus(self, contig):
        t = self.tables.get(contig)
        if t is None:
            return self.sparse.consensus(contig)
        counts = t["counts"].view(-1, N_LABELS)
        first = t["first"].view(-1, N_LABELS)
        best = counts.max(dim=1, keepdim=True).values
        voted = torch.nonzero(best[:, 0] > 0)[:, 0]
        cand = torch.where(counts[voted] == best[voted], first[voted], torch.full_like(first[voted], torch.iinfo(torch.int64).max))
        winner = cand.argmin(dim=1)
        keys = voted.cpu().numpy() + t["base"] * (MAX_INS + 1)
        return np.stack([keys // (MAX_INS + 1), keys % (MAX_INS + 1)], axis=1), winner.cpu().numpy()

    def release(self, contig):
        self.tables.pop(contig, None)
        self.sparse.tables.pop(contig, None)


class _SlabDataset(Dataset):
    """One item = up to ``chunk`` consecutive windows of one group, read as slabs (row f3 of SURVEY.md 8f).

    ``lo``/``hi`` restrict the dataset to the flat window range [lo, hi) (multi-GPU sharding: contiguous ranges
    keep labels aligned with the positions the stitcher needs, roko/inference.py:119-124); ``examples=False``
    reads positions only (rank 0 of a sharded run votes with labels gathered from the other ranks)."""

    def __init__(self, path, chunk, h5=None, lo=0, hi=None, examples=True):
        self.path, self.chunk, self._h5mod, self.fd, self.examples = path, chunk, h5, None, examples
        self.items, self.contigs, self.group_contig, self.total = [], {}, {}, 0
        fd = self._open()
        try:
            flat = 0
            for g in fd.keys():
                if g == "contigs":
                    continue
                n = int(fd[g].attrs["size"])
                self.group_contig[g] = fd[g].attrs["contig"]
                for a in range(0, n, chunk):
                    b = min(a + chunk, n)
                    a2, b2 = max(a, lo - flat), (b if hi is None else min(b, hi - flat))
                    if a2 < b2:
                        self.items.append((g, a2, b2, flat + a2))
                flat += n
            self.total = flat
            for k in fd["contigs"]:
                grp = fd["contigs"][k]
                self.contigs[str(k)] = (grp.attrs["seq"], grp.attrs["len"])
        finally:
            fd.close()

    def _open(self):
        return (self._h5mod or _h5(self.path)).File(self.path, "r")

    def __len__(self):
        return len(self.items)

    def __getitem__(self, i):
        if self.fd is None:
            self.fd = self._open()
        g, a, b, flat = self.items[i]
        grp = self.fd[g]
        x = torch.from_numpy(np.ascontiguousarray(grp["examples"][a:b])) if self.examples else torch.empty(0)
        return grp.attrs["contig"], torch.from_numpy(np.ascontiguousarray(grp["positions"][a:b])), x, flat

    def read_into(self, i, x_out):
        """Item i with its windows read STRAIGHT into ``x_out[:n]`` (a pinned staging tensor): ``Dataset.read_direct`` where the
        backend has it (h5py, the synthetic file) -- no intermediate array, no page-faulting allocation per slab -- else a copy.
        Returns (contig, positions, n, flat index of the first window)."""
        if self.fd is None:
            self.fd = self._open()
        g, a, b, flat = self.items[i]
        grp = self.fd[g]
        ex, n = grp["examples"], b - a
        dest = x_out.numpy()[:n]
        if hasattr(ex, "read_direct"):
            ex.read_direct(dest, np.s_[a:b])
        else:
            dest[...] = ex[a:b]
        return grp.attrs["contig"], torch.from_numpy(np.ascontiguousarray(grp["positions"][a:b])), n, flat


def _dist_env():
    import os
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def infer_fast(data, model_path, out, workers=0, batch_size=128, h5=None, device=None, chunk=8192, vote_budget=4 << 30,
               stats=None):
    """Throughput driver: slab reads -> predict_host (coalesced device passes) -> dense GPU vote -> stitch.

    Launched under ``torchrun`` (WORLD_SIZE > 1, one process per GPU) it shards the flat window range into contiguous
    per-rank ranges, broadcasts the weights from rank 0 over NCCL (one flat 4.4 MB broadcast instead of the
    reference's dormant per-forward ``nn.DataParallel`` replication, roko/inference.py:12,96-97), gathers the uint8
    labels (90 B / window) to rank 0 over NVLink, and rank 0 votes, stitches and writes the FASTA.  Returns the
    records on rank 0 and ``None`` elsewhere.  ``stats`` (a dict) receives wall-clock seconds per phase."""
    import time
    clock, t_last = {}, [time.perf_counter()]

    def lap(name):
        now = time.perf_counter()
        clock[name] = clock.get(name, 0.0) + now - t_last[0]
        t_last[0] = now

    if not torch.cuda.is_available():
        raise RuntimeError("roko_b200.inference needs a CUDA (sm_100a) device: the model path has no CPU fallback")
    rank, local_rank, world = _dist_env()
    if world > 1:
        import torch.distributed as dist
        from . import dist as rdist
        device = torch.device("cuda", local_rank)
        torch.cuda.set_device(device)
        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=device)
    else:
        device = torch.device(device or "cuda:0")
    model = RNN(IN_SIZE, HIDDEN_SIZE, NUM_LAYERS).to(device)
    if rank == 0:
        model.load_state_dict(torch.load(model_path, map_location=device))
    if world > 1:
        rdist.broadcast_weights(model, src=0)
    model.eval()
    lap("setup (process group, model, weights)")

    meta = _SlabDataset(data, chunk, h5=h5, examples=False)                # full index (positions only)
    lo, hi = (0, meta.total) if world == 1 else rdist.shard_range(meta.total, rank, world)
    dataset = _SlabDataset(data, chunk, h5=h5, lo=lo, hi=hi)
    # two pinned staging buffers: while the model path (a blocking C call that releases the GIL) runs on one from a worker thread,
    # the next slab is read straight into the other
    from concurrent.futures import ThreadPoolExecutor
    x_pin = [torch.empty((chunk, 200, 90), dtype=torch.uint8).pin_memory() for _ in range(2)]
    y_pin = [torch.empty((chunk, 90), dtype=torch.uint8).pin_memory() for _ in range(2)]
    votes = DenseVoteTable(device, budget_bytes=vote_budget) if rank == 0 else None
    remaining = {}                                                        # items still to vote per contig (rank 0)
    for g, a, b, _ in meta.items:
        remaining[meta.group_contig[g]] = remaining.get(meta.group_contig[g], 0) + 1
    records, order = {}, []

    def vote(contig, pos, labels):
        if contig not in remaining:
            return
        if contig not in order:
            order.append(contig)
        votes.add(contig, meta.contigs[contig][1], pos.reshape(-1, 2), labels.reshape(-1))
        remaining[contig] -= 1
        if remaining[contig] == 0:                                        # contig complete: stitch now, free its table
            positions, winners = votes.consensus(contig)
            records[contig] = stitch(meta.contigs[contig][0], positions, winners)
            votes.release(contig)

    if rank == 0:
        print("Inference started")
    local = torch.empty((hi - lo, 90), dtype=torch.uint8, device=device) if world > 1 else None
    done = 0
    lap("index")

    def run_model(slot, n):
        torch.cuda.set_device(device)
        model.predict_host(x_pin[slot][:n], batch=batch_size, out=y_pin[slot][:n], device=device)

    def finish(job):
        """Wait for the model call of a staged slab, then hand its labels to the gather buffer / the vote."""
        fut, slot, contig, pos, n, flat = job
        fut.result()
        lap("model path (H2D, kernels, D2H) not hidden behind reads")
        if world > 1:
            local[flat - lo:flat - lo + n].copy_(y_pin[slot][:n])
        else:
            vote(contig, pos, y_pin[slot][:n])
            lap("vote + stitch")

    loader = iter(DataLoader(dataset, batch_size=None, shuffle=False, num_workers=workers)) if workers else None
    with ThreadPoolExecutor(max_workers=1) as pool:
        pending = None
        for i in range(len(dataset)):
            slot = i & 1                                                  # the other slot may still be in flight (pending)
            if loader is None:
                contig, pos, n, flat = dataset.read_into(i, x_pin[slot])
            else:                                                         # worker processes read; one copy into pinned memory here
                contig, pos, x, flat = next(loader)
                n = x.shape[0]
                x_pin[slot][:n].copy_(x)
            lap("read + stage")
            job = (pool.submit(run_model, slot, n), slot, contig, pos, n, flat)
            if pending is not None:
                finish(pending)
            pending = job
            done += n
            if rank == 0 and (done // batch_size) % 100 == 0:
                print(f"{done // batch_size} batches processed")
        if pending is not None:
            finish(pending)
    model.check_codes()                                                   # nn.Embedding would have raised on a bad code
    if world > 1:
        labels = rdist.gather_labels(local, meta.total)                   # (N, 90) uint8 on rank 0, rank order = file order
        torch.cuda.synchronize(device)
        lap("label gather (NCCL)")
        if rank != 0:
            if stats is not None:
                stats.update(clock)
            return None
        for i in range(len(meta)):
            contig, pos, _, flat = meta[i]
            vote(contig, pos, labels[flat:flat + pos.shape[0]])         # labels stay on the GPU: the vote tables live there
        lap("vote + stitch")
    out_records = [(c, records[c]) for c in order]
    write_fasta(out_records, out)
    lap("fasta")
    if stats is not None:
        stats.update(clock)
    return out_records


def main():
    parser = argparse.ArgumentParser()
    parser.add_argument("data", type=str)
    parser.add_argument("model", type=str)
    parser.add_argument("out", type=str)
    parser.add_argument("--t", type=int, default=0)
    parser.add_argument("--b", type=int, default=128)
    parser.add_argument("--per-batch", action="store_true", help="the reference's per-batch loop instead of slab reads")
    args = parser.parse_args()
    (infer if args.per_batch else infer_fast)(args.data, args.model, args.out, args.t, args.b)


if __name__ == "__main__":
    main()
