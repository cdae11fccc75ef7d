"""CPU: the synthetic feature file (roko_b200/synth.py) honours the roko .hdf5 schema and is reproducible by range."""
import numpy as np

from roko_b200 import inference as inf
from roko_b200 import synth


def test_synthetic_file_schema_and_reproducibility():
    path = "synthetic://9000?contig_len=120000&group=700&seed=5"
    f = synth.File(path)
    groups = [k for k in f.keys() if k != "contigs"]
    assert sum(int(f[g].attrs["size"]) for g in groups) == 9000
    names = list(f["contigs"].keys())
    assert all(len(f["contigs"][n].attrs["seq"]) == f["contigs"][n].attrs["len"] == 120000 for n in names)
    g = f[groups[3]]
    x = g["examples"][10:20]
    assert x.shape == (10, 200, 90) and x.dtype == np.uint8 and x.max() < 12
    assert np.array_equal(synth.File(path)[groups[3]]["examples"][12:15], x[2:5])       # any sub-range reproduces
    pos = g["positions"][0:2]
    assert pos.shape == (2, 90, 2) and (pos[..., 1] == 0).all() and (pos[1, :, 0] - pos[0, :, 0] == 30).all()
    assert int(pos[..., 0].max()) < 120000
    ds = inf._SlabDataset(path, 512, h5=synth)
    assert ds.total == 9000 and sum(b - a for _, a, b, _ in ds.items) == 9000
    c, p, xs, flat = ds[len(ds) - 1]
    assert xs.shape[0] == p.shape[0] and flat + xs.shape[0] == 9000


def test_read_direct_and_read_into_match_slicing():
    import torch
    path = "synthetic://5000?contig_len=90000&group=800&seed=3"
    ds = inf._SlabDataset(path, 300, h5=synth, lo=450, hi=4100)
    buf = torch.zeros((300, 200, 90), dtype=torch.uint8)
    flat_expected = 450
    for i in range(len(ds)):
        c, p, x, flat = ds[i]
        c2, p2, n, flat2 = ds.read_into(i, buf)
        assert (c, flat, n) == (c2, flat2, x.shape[0]) and flat == flat_expected
        assert torch.equal(buf[:n], x) and torch.equal(p, p2)
        flat_expected += n
    assert flat_expected == 4100
