"""In-memory stand-in for the tiny part of h5py the roko feature format uses (h5py is not in this
image): File(path, 'r') -> groups with .attrs and datasets indexable by row."""
import numpy as np

_FILES = {}


class _Group(dict):
    def __init__(self, attrs=None, **datasets):
        super().__init__(**datasets)
        self.attrs = dict(attrs or {})


class File:
    def __init__(self, path, mode="r", **kw):
        self._root = _FILES[path]

    def keys(self):
        return list(self._root.keys())

    def __getitem__(self, k):
        return self._root[k]

    def close(self):
        pass


def register(path, contigs, groups):
    """contigs: {name: seq};  groups: [(group_name, contig, positions (N,90,2) i64, examples (N,200,90) u8
    [, labels (N,90)])] -- the 5-tuple form is a training file (reference data.py:40-48)."""
    root = {"contigs": _Group()}
    for name, seq in contigs.items():
        root["contigs"][name] = _Group({"name": name, "seq": seq, "len": len(seq)})
    for gname, contig, pos, ex, *rest in groups:
        root[gname] = _Group({"contig": contig, "size": len(ex)}, positions=np.asarray(pos, np.int64),
                             examples=np.asarray(ex, np.uint8))
        if rest:
            root[gname]["labels"] = np.asarray(rest[0], np.int64)
    _FILES[path] = root
