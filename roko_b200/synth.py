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
