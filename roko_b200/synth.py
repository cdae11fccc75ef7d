"""Seeded synthetic inputs for the roko hot path (there is no network / dataset here).

Two generators, both numpy-PCG64 so the bytes do not depend on the torch build:

* ``uniform_windows``    -- uniform codes 0..11; run time of the path is data independent, so
                            this is what throughput is measured on.
* ``structured_windows`` -- pileup-like windows used for parity: a per-window truth row shared
                            by the 200 sampled reads, 10 % substituted cells, 20 % of the reads
                            with an UNKNOWN (code 5) prefix as produced at read ends
                            (reference generate.cpp:134-136) and a per-read strand offset of +6
                            (reference generate.cpp:145).  Gives mixed labels and near-ties.
"""
import numpy as np

READS, COLS, N_CODES = 200, 90, 12


def uniform_windows(n, seed=1234):
    rng = np.random.Generator(np.random.PCG64(seed))
    return rng.integers(0, N_CODES, size=(n, READS, COLS), dtype=np.uint8)


def structured_windows(n, seed=101, return_truth=False):
    rng = np.random.Generator(np.random.PCG64(seed))
    truth = rng.integers(0, 5, size=(n, 1, COLS), dtype=np.uint8)
    x = np.broadcast_to(truth, (n, READS, COLS)).copy()
    noise = rng.random((n, READS, COLS)) < 0.10
    x[noise] = rng.integers(0, 5, size=int(noise.sum()), dtype=np.uint8)
    clipped = rng.random((n, READS)) < 0.20
    plen = rng.integers(0, COLS + 1, size=(n, READS))
    prefix = (np.arange(COLS)[None, None, :] < plen[:, :, None]) & clipped[:, :, None]
    x[prefix] = 5
    strand = rng.random((n, READS, 1)) < 0.5
    x = x + (6 * strand).astype(np.uint8)
    x = np.ascontiguousarray(x, dtype=np.uint8)
    return (x, truth[:, 0, :].copy()) if return_truth else x


# ---- synthetic feature FILE: an h5py-like stand-in for BASELINE.json configs 3-5 ---------------------------
# ``File("synthetic://<n_windows>[?contig_len=..&group=..&seed=..]")`` exposes the roko feature-file schema
# (SURVEY.md App. C / reference roko/data.py:40-48): one group per region with ``attrs['contig'|'size']``,
# ``positions`` (n, 90, 2) int64 and ``examples`` (n, 200, 90) uint8, plus ``/contigs/<name>`` with the draft.
# Windows are generated on demand from the seed (slab reads of any range cost only that range), so a
# 1 M-window "file" (18 GB of examples) needs no disk or RAM: ``inference.infer_fast(..., h5=synth)``.
_BLOCK_CACHE = {}


class _LazyRows:
    def __init__(self, n, make, make_into=None):
        self.n, self.make, self.make_into = n, make, make_into
        self.shape = (n,)
        if make_into is not None:
            self.read_direct = self._read_direct          # h5py.Dataset.read_direct(dest, source_sel): fill a caller-owned array

    def _read_direct(self, dest, source_sel=None, dest_sel=None):
        a, b, step = (source_sel if isinstance(source_sel, slice) else slice(None)).indices(self.n)
        assert step == 1 and dest_sel is None
        self.make_into(a, b, dest)

    def __len__(self):
        return self.n

    def __getitem__(self, idx):
        if isinstance(idx, slice):
            a, b, step = idx.indices(self.n)
            assert step == 1
            return self.make(a, b)
        return self.make(int(idx), int(idx) + 1)[0]


class _SynGroup(dict):
    def __init__(self, attrs, **datasets):
        super().__init__(**datasets)
        self.attrs = attrs


_FILES = {}


def File(path, mode="r", **kw):
    """h5py.File stand-in: files are immutable functions of their path, so one object per process serves every open."""
    f = _FILES.get(path)
    if f is None:
        f = _FILES[path] = _File(path)
    return f


class _File:
    WINDOW_STEP = 30            # windows slide by 30 positions (reference include/generate.h:21)

    def __init__(self, path, mode="r", **kw):
        assert path.startswith("synthetic://"), path
        spec, _, query = path[len("synthetic://"):].partition("?")
        opts = dict(kv.split("=") for kv in query.split("&") if kv)
        self.n = int(spec)
        self.seed = int(opts.get("seed", 1234))
        self.contig_len = int(opts.get("contig_len", 3_000_000))
        self.group = int(opts.get("group", 3300))          # windows per group (a 100 kb region holds ~3 300)
        self.cache = int(opts.get("cache", 0))
        self.labels = int(opts.get("labels", 0))           # labelled (training) file: every group also has ``labels`` (n, 90) int64 in 0..4
        per_contig = max(1, (self.contig_len - 90) // self.WINDOW_STEP)
        self._root = {"contigs": _SynGroup({})}
        first, ci = 0, 0
        while first < self.n:
            cn = min(per_contig, self.n - first)
            name = f"ctg{ci}"
            rng = np.random.Generator(np.random.PCG64(self.seed + 7919 * ci))
            draft = "".join(np.array(list("ACGT"))[rng.integers(0, 4, size=self.contig_len)])
            self._root["contigs"][name] = _SynGroup({"name": name, "seq": draft, "len": self.contig_len})
            for g0 in range(0, cn, self.group):
                gn = min(self.group, cn - g0)
                self._root[f"{name}_{g0}"] = _SynGroup(
                    {"contig": name, "size": gn},
                    positions=_LazyRows(gn, lambda a, b, w0=g0: self._positions(w0 + a, w0 + b)),
                    examples=_LazyRows(gn, lambda a, b, w0=first + g0: self._examples(w0 + a, w0 + b),
                                       lambda a, b, out, w0=first + g0: self._examples(w0 + a, w0 + b, out)))
                if self.labels:
                    self._root[f"{name}_{g0}"]["labels"] = _LazyRows(gn, lambda a, b, w0=first + g0: self._labels(w0 + a, w0 + b))
            first += cn
            ci += 1

    def _positions(self, a, b):
        """Window i of a contig covers reference positions 30 i .. 30 i + 89 (no insertion slots)."""
        start = (np.arange(a, b, dtype=np.int64) * self.WINDOW_STEP)[:, None] + np.arange(COLS, dtype=np.int64)[None, :]
        return np.stack([start, np.zeros_like(start)], axis=2)

    def _examples(self, a, b, out=None):
        if out is None:
            out = np.empty((b - a, READS, COLS), dtype=np.uint8)
        blk = 4096                                             # generated in seed-addressed blocks: any range is reproducible
        for k in range(a // blk, (b - 1) // blk + 1):
            lo, hi = max(a, k * blk), min(b, (k + 1) * blk)
            block = self._block(k, min(blk, self.n - k * blk))
            out[lo - a:hi - a] = block[lo - k * blk:hi - k * blk]
        return out

    def _labels(self, a, b):
        """Uniform labels 0..4 per (window, column), reproducible by range like the examples."""
        out = np.empty((b - a, COLS), dtype=np.int64)
        blk = 65536
        for k in range(a // blk, (b - 1) // blk + 1):
            lo, hi = max(a, k * blk), min(b, (k + 1) * blk)
            rng = np.random.Generator(np.random.PCG64(self.seed * 7_000_003 + k))
            block = rng.integers(0, 5, size=(min(blk, self.n - k * blk), COLS), dtype=np.int64)
            out[lo - a:hi - a] = block[lo - k * blk:hi - k * blk]
        return out

    def _block(self, k, rows):
        """Block k of the examples (4 096 windows), drawn whole so that any sub-range is reproducible: 64-bit draws viewed as
        bytes and folded into 0..11 (the slight non-uniformity, 256 = 21 x 12 + 4, is irrelevant for a throughput workload).
        With ``?cache=1`` blocks are kept (per process), so a second pass over the file reads memory, like a page-cached .hdf5."""
        key = (self.seed, self.n, k)
        if self.cache and key in _BLOCK_CACHE:
            return _BLOCK_CACHE[key]
        rng = np.random.Generator(np.random.PCG64(self.seed * 1_000_003 + k))
        raw = rng.integers(0, 2 ** 63, size=(rows * READS * COLS + 7) // 8, dtype=np.uint64).view(np.uint8)[:rows * READS * COLS]
        block = (raw % N_CODES).reshape(rows, READS, COLS)
        if self.cache:
            _BLOCK_CACHE[key] = block
        return block

    def keys(self):
        return list(self._root.keys())

    def __getitem__(self, k):
        return self._root[k]

    def close(self):
        pass
