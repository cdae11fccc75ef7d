"""Host side of the inference entry point: vote table, stitching and FASTA writer against a plain
restatement of the reference's loops (inference.py:101,119-154); dataset index against the format."""
import itertools
from collections import Counter, defaultdict

import numpy as np
import pytest

from roko_b200 import inference as inf
from tests import fake_h5


def reference_vote_and_stitch(contigs, batches):
    """Restates inference.py:101,119-151 (Counter per position, most_common tie rule, stitching)."""
    result = defaultdict(lambda: defaultdict(lambda: Counter()))
    for c, pos, Y in batches:
        for cb, pb, yb in zip(c, pos, Y):
            for p, y in zip(pb, yb):
                result[cb][(int(p[0]), int(p[1]))][inf.decoding[int(y)]] += 1
    out = []
    for contig in result:
        values = result[contig]
        pos_sorted = sorted(values)
        pos_sorted = list(itertools.dropwhile(lambda x: x[1] != 0, pos_sorted))
        first = pos_sorted[0][0]
        seq = contigs[contig][:first]
        for p in pos_sorted:
            base, _ = values[p].most_common(1)[0]
            if base == inf.GAP:
                continue
            seq += base
        seq += contigs[contig][pos_sorted[-1][0] + 1:]
        out.append((contig, seq))
    return out


def make_windows(rng, contig_len, n_windows):
    """Overlapping 90-slot windows sliding by 30 (generate.cpp:152-155) with a few insertion slots."""
    slots = []
    for r in range(5, contig_len - 5):
        slots.append((r, 0))
        if rng.random() < 0.15:
            slots.append((r, 1))
    slots = np.array(slots, np.int64)
    pos = np.stack([slots[30 * i:30 * i + 90] for i in range(n_windows) if 30 * i + 90 <= len(slots)])
    return pos


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_vote_table_matches_counter_semantics(seed):
    rng = np.random.default_rng(seed)
    contigs = {"ctgA": "".join(rng.choice(list("ACGT"), 400)), "ctgB": "".join(rng.choice(list("ACGT"), 300))}
    batches, votes = [], inf.VoteTable()
    for name in contigs:
        pos = make_windows(rng, len(contigs[name]), 9)
        # noisy labels so that 1-1 and 1-1-1 ties occur between overlapping windows
        Y = rng.integers(0, 5, size=pos.shape[:2]).astype(np.uint8)
        for b0 in range(0, len(pos), 4):
            sl = slice(b0, b0 + 4)
            batches.append(([name] * len(pos[sl]), pos[sl], Y[sl]))
    for c, pos, Y in batches:
        votes.add(c[0], pos.reshape(-1, 2), Y.reshape(-1))
    mine = []
    for contig in votes.tables:
        p, w = votes.consensus(contig)
        mine.append((contig, inf.stitch(contigs[contig], p, w)))
    assert mine == reference_vote_and_stitch(contigs, batches)


@pytest.mark.parametrize("seed", [3, 4])
def test_dense_vote_table_matches_counter_semantics(seed):
    import torch
    rng = np.random.default_rng(seed)
    contigs = {"ctgA": "".join(rng.choice(list("ACGT"), 500))}
    pos = make_windows(rng, 500, 12)
    Y = rng.integers(0, 5, size=pos.shape[:2]).astype(np.uint8)
    batches, votes = [], inf.DenseVoteTable("cpu")
    for b0 in range(0, len(pos), 5):
        sl = slice(b0, b0 + 5)
        batches.append((["ctgA"] * len(pos[sl]), pos[sl], Y[sl]))
        votes.add("ctgA", 500, torch.from_numpy(pos[sl].reshape(-1, 2)), torch.from_numpy(Y[sl].reshape(-1)))
    p, w = votes.consensus("ctgA")
    assert [("ctgA", inf.stitch(contigs["ctgA"], p, w))] == reference_vote_and_stitch(contigs, batches)


@pytest.mark.parametrize("budget,reverse", [(4 << 30, False), (4 << 30, True), (20_000, False), (0, True)])
def test_dense_vote_table_grows_spills_and_releases(budget, reverse):
    """Tables are sized from the positions seen (growing in either direction), fall back to the sparse table over
    budget (mid-contig too), and are freed per contig -- with Counter semantics intact throughout."""
    import torch
    rng = np.random.default_rng(7)
    contigs = {"ctgA": "".join(rng.choice(list("ACGT"), 900)), "ctgB": "".join(rng.choice(list("ACGT"), 300))}
    votes, batches = inf.DenseVoteTable("cpu", budget_bytes=budget), []
    for name in contigs:
        pos = make_windows(rng, len(contigs[name]), 25)
        Y = rng.integers(0, 5, size=pos.shape[:2]).astype(np.uint8)
        starts = list(range(0, len(pos), 4))
        for b0 in (reversed(starts) if reverse else starts):
            sl = slice(b0, b0 + 4)
            batches.append(([name] * len(pos[sl]), pos[sl], Y[sl]))
            votes.add(name, 10, torch.from_numpy(pos[sl].reshape(-1, 2)), torch.from_numpy(Y[sl].reshape(-1)))   # wrong attrs['len'] is harmless
    assert votes.contigs() == ["ctgA", "ctgB"]
    mine = []
    for contig in votes.contigs():
        p, w = votes.consensus(contig)
        mine.append((contig, inf.stitch(contigs[contig], p, w)))
        votes.release(contig)
    assert mine == reference_vote_and_stitch(contigs, batches)
    assert not votes.tables and not votes.sparse.tables
    with pytest.raises(IndexError):
        votes.add("ctgA", 900, torch.tensor([[5, 4]]), torch.tensor([1], dtype=torch.uint8))      # ins > MAX_INS


def test_slab_dataset_shards_cover_the_file_once():
    rng = np.random.default_rng(5)
    ex = rng.integers(0, 12, (11, 200, 90), dtype=np.uint8)
    pos = rng.integers(0, 50, (11, 90, 2)).astype(np.int64)
    fake_h5.register("mem://shard", {"c": "ACGT" * 50}, [("g0", "c", pos[:4], ex[:4]), ("g1", "c", pos[4:9], ex[4:9]), ("g2", "c", pos[9:], ex[9:])])
    from roko_b200.dist import shard_range
    for world in (1, 2, 3, 4):
        seen = []
        for r in range(world):
            lo, hi = shard_range(11, r, world)
            ds = inf._SlabDataset("mem://shard", 3, h5=fake_h5, lo=lo, hi=hi)
            for i in range(len(ds)):
                c, p, x, flat = ds[i]
                assert np.array_equal(x.numpy(), ex[flat:flat + len(x)]) and np.array_equal(p.numpy(), pos[flat:flat + len(x)])
                seen.extend(range(flat, flat + len(x)))
        assert seen == list(range(11))


def test_fasta_writer_format(tmp_path):
    path = tmp_path / "o.fasta"
    inf.write_fasta([("c1", "A" * 130), ("c2", "ACGT")], str(path))
    assert path.read_text() == (">c1 <unknown description>\n" + "A" * 60 + "\n" + "A" * 60 + "\n" + "A" * 10 + "\n"
                                ">c2 <unknown description>\nACGT\n")


def test_dataset_index_and_items():
    rng = np.random.default_rng(3)
    ex = rng.integers(0, 12, (5, 200, 90), dtype=np.uint8)
    pos = np.zeros((5, 90, 2), np.int64)
    fake_h5.register("mem://t1", {"ctg": "ACGT" * 50}, [("ctg_0-100", "ctg", pos[:3], ex[:3]), ("ctg_100-200", "ctg", pos[3:], ex[3:])])
    ds = inf.InferenceDataset("mem://t1", transform=inf.ToTensor(), h5=fake_h5)
    assert len(ds) == 5 and ds.contigs["ctg"] == ("ACGT" * 50, 200)
    c, p, x = ds[4]
    assert c == "ctg" and x.dtype.is_floating_point is False and np.array_equal(x.numpy(), ex[4])


@pytest.mark.gpu
def test_infer_end_to_end_on_gpu(tmp_path, seed1_state, seed1_weights):
    """inference entry point on the GPU == oracle labels pushed through the reference's vote/stitch."""
    import torch
    from oracle import roko_oracle as O
    from roko_b200.synth import structured_windows
    rng = np.random.default_rng(11)
    draft = "".join(rng.choice(list("ACGT"), 700))
    pos = make_windows(rng, len(draft), 20)
    x = structured_windows(len(pos), seed=321)
    fake_h5.register("mem://e2e", {"ctg1": draft}, [("ctg1_0-350", "ctg1", pos[:11], x[:11]), ("ctg1_350-700", "ctg1", pos[11:], x[11:])])
    pth = tmp_path / "m.pth"
    torch.save(seed1_state, pth)
    recs = inf.infer("mem://e2e", str(pth), str(tmp_path / "out.fasta"), workers=0, batch_size=7, h5=fake_h5)
    labels = O.predict(x, seed1_weights)
    batches = [(["ctg1"] * len(pos[i:i + 7]), pos[i:i + 7], labels[i:i + 7]) for i in range(0, len(pos), 7)]
    assert recs == reference_vote_and_stitch({"ctg1": draft}, batches)
    assert (tmp_path / "out.fasta").read_text().startswith(">ctg1 <unknown description>\n")
    # the slab / predict_host / dense-vote driver gives the same consensus (votes arrive group by group, in order)
    fast = inf.infer_fast("mem://e2e", str(pth), str(tmp_path / "out_fast.fasta"), workers=0, batch_size=7, h5=fake_h5, chunk=6)
    assert fast == recs
    assert (tmp_path / "out_fast.fasta").read_text() == (tmp_path / "out.fasta").read_text()
