"""CPU oracle for roko's inference hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain-numpy restatement of ``RNN.forward`` (reference roko/rnn_model.py:46-59) plus the
argmax the caller applies (reference roko/inference.py:116).  Only ``tests/``,
``__graft_entry__.smoke()`` and the ``cpu_baseline`` / ``--impl reference`` legs of
``bench.py`` may import this module; the shipped path (``roko_b200``) never does and
fails loudly when its CUDA library is missing.

Parity pin: the reference ships no tests or golden vectors for this path (SURVEY.md
section 8c), so the oracle is pinned against outputs of the reference class itself,
executed in the build container by ``oracle/make_golden.py`` (imports
/root/reference/roko/rnn_model.py, seeded weights + seeded structured pileups) and
committed under ``tests/golden/``.  ``tests/test_oracle.py`` checks this module against
those fixtures.

Weights are passed as a dict ``name -> ndarray`` with exactly the 31 state_dict keys of
the reference module (rnn_model.py:25-44).

The arithmetic lives in PyTorch (torch==1.3.1 pinned by the reference's
requirements.txt:8; not vendored).  Restated semantics:
  * nn.Embedding  : row gather                                      (rnn_model.py:28,47)
  * nn.Linear     : y = x W^T + b                                   (rnn_model.py:31,34,44)
  * nn.GRU        : gate order [r; z; n], b_hn inside r*(...),
                    h0 = 0, reverse direction written at its own t   (rnn_model.py:40-41,57)
  * dropout       : identity in eval mode                           (rnn_model.py:29,32,35)
"""
import numpy as np

READS = 200      # rows of a window: sampled reads        (reference include/generate.h:19)
COLS = 90        # columns of a window: pileup positions   (reference include/generate.h:19)
N_CODES = 12     # embedding rows                         (rnn_model.py:28)
EMB = 50         # embedding dim                          (rnn_model.py:28)
FC1 = 100        # rnn_model.py:31
FC2 = 10         # rnn_model.py:34
IN_SIZE = 500    # rnn_model.py:10
HIDDEN = 128     # rnn_model.py:11
LAYERS = 3       # rnn_model.py:12
CLASSES = 5      # rnn_model.py:44

STATE_KEYS = (
    ["embedding.weight", "fc1.weight", "fc1.bias", "fc2.weight", "fc2.bias"]
    + [f"gru.{kind}_l{l}{sfx}"
       for l in range(LAYERS) for sfx in ("", "_reverse")
       for kind in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
    + ["fc4.weight", "fc4.bias"]
)

STATE_SHAPES = {
    "embedding.weight": (N_CODES, EMB),
    "fc1.weight": (FC1, READS), "fc1.bias": (FC1,),
    "fc2.weight": (FC2, FC1), "fc2.bias": (FC2,),
    "fc4.weight": (CLASSES, 2 * HIDDEN), "fc4.bias": (CLASSES,),
}
for _l in range(LAYERS):
    for _s in ("", "_reverse"):
        STATE_SHAPES[f"gru.weight_ih_l{_l}{_s}"] = (3 * HIDDEN, IN_SIZE if _l == 0 else 2 * HIDDEN)
        STATE_SHAPES[f"gru.weight_hh_l{_l}{_s}"] = (3 * HIDDEN, HIDDEN)
        STATE_SHAPES[f"gru.bias_ih_l{_l}{_s}"] = (3 * HIDDEN,)
        STATE_SHAPES[f"gru.bias_hh_l{_l}{_s}"] = (3 * HIDDEN,)


def _w(weights, key, dtype):
    a = np.asarray(weights[key])
    assert a.shape == STATE_SHAPES[key], (key, a.shape)
    return a.astype(dtype, copy=False)


def _sigmoid(v):
    return 1.0 / (1.0 + np.exp(-v))


def check_input(x):
    """Shape / value-domain contract of the path (codes 0..11: reference generate.cpp:18-25,145)."""
    x = np.asarray(x)
    if x.ndim != 3 or x.shape[1] != READS or x.shape[2] != COLS:
        raise ValueError(f"expected (B,{READS},{COLS}), got {x.shape}")
    if x.size and (x.min() < 0 or x.max() >= N_CODES):
        raise IndexError("pileup code out of range 0..11")   # nn.Embedding raises IndexError on CPU
    return x


def front_end(x, weights, dtype=np.float32):
    """rnn_model.py:47-56 -- embedding, read-axis MLP, flatten to (B, 90, 500).

    Written as the reference computes it (gather, permute to [b, col, emb, read], contract
    the read axis), not with the one-hot factorisation the CUDA kernel uses.
    """
    x = check_input(x)
    E = _w(weights, "embedding.weight", dtype)
    W1, b1 = _w(weights, "fc1.weight", dtype), _w(weights, "fc1.bias", dtype)
    W2, b2 = _w(weights, "fc2.weight", dtype), _w(weights, "fc2.bias", dtype)
    B = x.shape[0]
    out = np.empty((B, COLS, IN_SIZE), dtype=dtype)
    W1T = np.ascontiguousarray(W1.T)
    W2T = np.ascontiguousarray(W2.T)
    for b in range(B):                                   # per window keeps the 3.6 MB gather cache-sized
        e = E[x[b].astype(np.int64)]                     # (200, 90, 50)   rnn_model.py:47
        e = np.ascontiguousarray(e.transpose(1, 2, 0))   # (90, 50, 200)   rnn_model.py:48
        a = np.maximum(e.reshape(-1, READS) @ W1T + b1, 0)   # (4500, 100)  rnn_model.py:50
        g = np.maximum(a @ W2T + b2, 0)                  # (4500, 10)      rnn_model.py:53
        out[b] = g.reshape(COLS, IN_SIZE)                # f = 10*e + k    rnn_model.py:56
    return out


def gru_direction(v, W_ih, W_hh, b_ih, b_hh, reverse):
    """One direction of one nn.GRU layer over v (B, T, in) -> (B, T, H); h0 = 0."""
    B, T, _ = v.shape
    H = W_hh.shape[1]
    gi = v @ W_ih.T + b_ih                               # (B, T, 3H), rows ordered [r; z; n]
    W_hhT = np.ascontiguousarray(W_hh.T)
    h = np.zeros((B, H), dtype=v.dtype)
    out = np.empty((B, T, H), dtype=v.dtype)
    steps = range(T - 1, -1, -1) if reverse else range(T)
    for t in steps:
        gh = h @ W_hhT + b_hh
        r = _sigmoid(gi[:, t, :H] + gh[:, :H])
        z = _sigmoid(gi[:, t, H:2 * H] + gh[:, H:2 * H])
        n = np.tanh(gi[:, t, 2 * H:] + r * gh[:, 2 * H:])     # b_hn is inside r*( . )
        h = (1 - z) * n + z * h
        out[:, t] = h
    return out


def gru(u, weights, dtype=np.float32, taps=None):
    """rnn_model.py:57 -- 3-layer bidirectional GRU, batch_first, eval mode."""
    v = u
    for l in range(LAYERS):
        halves = []
        for sfx, rev in (("", False), ("_reverse", True)):
            halves.append(gru_direction(
                v,
                _w(weights, f"gru.weight_ih_l{l}{sfx}", dtype), _w(weights, f"gru.weight_hh_l{l}{sfx}", dtype),
                _w(weights, f"gru.bias_ih_l{l}{sfx}", dtype), _w(weights, f"gru.bias_hh_l{l}{sfx}", dtype), rev))
        v = np.concatenate(halves, axis=2)               # [fwd ; bwd]
        if taps is not None:
            taps[f"gru_l{l}"] = v
    return v


def head(h, weights, dtype=np.float32):
    """rnn_model.py:59 -- fc4."""
    return h @ _w(weights, "fc4.weight", dtype).T + _w(weights, "fc4.bias", dtype)


def forward(x, weights, dtype=np.float32, taps=None):
    """RNN.forward (rnn_model.py:46-59): x (B,200,90) integer codes -> logits (B,90,5)."""
    u = front_end(x, weights, dtype)
    if taps is not None:
        taps["front"] = u
    h = gru(u, weights, dtype, taps)
    return head(h, weights, dtype)


def labels_from_logits(logits):
    """inference.py:116 -- torch.argmax(logits, dim=2): first index of the maximum."""
    return np.argmax(logits, axis=2).astype(np.uint8)


def top2_gap(logits):
    s = np.sort(logits, axis=2)
    return s[..., -1] - s[..., -2]


def predict(x, weights, dtype=np.float32):
    return labels_from_logits(forward(x, weights, dtype))
