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
