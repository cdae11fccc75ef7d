"""CPU oracle for the TRAINING use of roko's network  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Restates the train-mode forward of the reference module (roko/rnn_model.py:46-59 with the dropout
sites :29 embedding, :32 after fc1, :35 after fc2 and nn.GRU's inter-layer dropout :41) as plain
torch tensor algebra in float64 on the CPU, with the dropout keep-masks passed IN instead of
drawn, and lets ``torch.autograd`` differentiate it -- the same engine the reference's training
loop relies on (roko/train.py:46-53: ``F.cross_entropy(model(x).transpose(1, 2), y)``; backward).

With every mask absent it is the eval-mode function; ``oracle/make_train_golden.py`` pins it there
against the gradients of the reference class itself (tests/golden/train_seed1.npz).  With masks it
is the only way to check a kernel's dropout bit for bit: the kernels export their masks
(``roko_b200_dropout_mask``) and this module applies the same ones.

Only ``tests/`` may import this module.
"""
import numpy as np
import torch
import torch.nn.functional as F

from .roko_oracle import CLASSES, COLS, HIDDEN, LAYERS, READS, STATE_KEYS  # noqa: F401

SITES = ("emb", "fc1", "fc2", "gru0", "gru1")          # site numbers 0..4 of roko_b200_dropout_mask


def mask_shapes(batch):
    return {"emb": (batch, READS, COLS, 50), "fc1": (batch, COLS, 50, 100), "fc2": (batch, COLS, 50, 10),
            "gru0": (batch, COLS, 2 * HIDDEN), "gru1": (batch, COLS, 2 * HIDDEN)}


def kernel_keep_masks(p, seed, batch):
    """numpy restatement of the counter-based masks of roko_b200/csrc/train.cuh (drop_hash / drop_keep):
    (k1, k2) = hi32, lo32 of splitmix64(seed ^ (s+1) * K);  keep element i of site s  <=>  p == 0  or
    lowbias32(i + k1) ^ k2 >= floor(p * 2^32)   (32-bit multiply-xorshift finaliser, arithmetic mod 2^32)."""
    out = {}
    thresh = np.uint32(int(p * 4294967296.0))
    with np.errstate(over="ignore"):
        for site, (name, shape) in enumerate(mask_shapes(batch).items()):
            n = int(np.prod(shape))
            if thresh == 0:
                out[name] = np.ones(shape, dtype=np.uint8)
                continue
            z = np.uint64(seed) ^ (np.uint64(site + 1) * np.uint64(0xD1B54A32D192ED03))
            z = z + np.uint64(0x9E3779B97F4A7C15)
            z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
            z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
            z = z ^ (z >> np.uint64(31))
            k1, k2 = np.uint32(z >> np.uint64(32)), np.uint32(z & np.uint64(0xFFFFFFFF))
            x = np.arange(n, dtype=np.uint64).astype(np.uint32) + k1
            x = (x ^ (x >> np.uint32(16))) * np.uint32(0x7FEB352D)
            x = (x ^ (x >> np.uint32(15))) * np.uint32(0x846CA68B)
            x = x ^ (x >> np.uint32(16))
            out[name] = ((x ^ k2) >= thresh).astype(np.uint8).reshape(shape)
    return out


def _drop(t, masks, site, p):
    if masks is None or site not in masks or masks[site] is None:
        return t
    m = torch.as_tensor(np.asarray(masks[site]), dtype=t.dtype).reshape(t.shape)
    return t * m / (1.0 - p)                                # nn.Dropout: zero with prob p, scale the rest


def _gru_direction(seq, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of one nn.GRU layer, h0 = 0 (gate order r, z, n; b_hn inside r * (.))."""
    batch, steps, _ = seq.shape
    gi_all = seq @ w_ih.T + b_ih
    h = torch.zeros((batch, HIDDEN), dtype=seq.dtype)
    outs = [None] * steps
    order = range(steps - 1, -1, -1) if reverse else range(steps)
    for t in order:
        gi = gi_all[:, t]
        gh = h @ w_hh.T + b_hh
        r = torch.sigmoid(gi[:, :HIDDEN] + gh[:, :HIDDEN])
        z = torch.sigmoid(gi[:, HIDDEN:2 * HIDDEN] + gh[:, HIDDEN:2 * HIDDEN])
        n = torch.tanh(gi[:, 2 * HIDDEN:] + r * gh[:, 2 * HIDDEN:])
        h = (1.0 - z) * n + z * h
        outs[t] = h
    return torch.stack(outs, dim=1)


def forward(params, x, masks=None, p=0.2):
    """params: dict name -> float64 tensor (31 state_dict keys); x: (B,200,90) integer codes."""
    xt = torch.as_tensor(np.asarray(x)).long()
    e = params["embedding.weight"][xt]                                        # rnn_model.py:47
    e = _drop(e, masks, "emb", p)
    e = e.permute(0, 2, 3, 1)                                                 # :48
    a = torch.relu(e @ params["fc1.weight"].T + params["fc1.bias"])           # :50
    a = _drop(a, masks, "fc1", p)                                             # :51
    g = torch.relu(a @ params["fc2.weight"].T + params["fc2.bias"])           # :53
    g = _drop(g, masks, "fc2", p)                                             # :54
    h = g.reshape(-1, COLS, 500)                                              # :56
    for layer in range(LAYERS):                                               # :57
        halves = []
        for sfx, rev in (("", False), ("_reverse", True)):
            halves.append(_gru_direction(h, params[f"gru.weight_ih_l{layer}{sfx}"],
                                         params[f"gru.weight_hh_l{layer}{sfx}"],
                                         params[f"gru.bias_ih_l{layer}{sfx}"],
                                         params[f"gru.bias_hh_l{layer}{sfx}"], rev))
        h = torch.cat(halves, dim=2)
        if layer + 1 < LAYERS:
            h = _drop(h, masks, f"gru{layer}", p)
    return h @ params["fc4.weight"].T + params["fc4.bias"]                    # :59


def relu_margin(state, x, masks=None, p=0.2):
    """Smallest |pre-activation| of the two ReLUs (rnn_model.py:50,53) over the batch, float64.

    ReLU's derivative jumps at 0: an fp32 implementation and this float64 one can disagree on the side
    of a pre-activation closer to 0 than fp32 rounding noise (~1e-6), which moves a few gradient entries
    by far more than rounding.  Parity tests use inputs whose margin is clear of that."""
    with torch.no_grad():
        prm = {k: torch.tensor(np.asarray(state[k]), dtype=torch.float64) for k in STATE_KEYS}
        e = _drop(prm["embedding.weight"][torch.as_tensor(np.asarray(x)).long()], masks, "emb", p).permute(0, 2, 3, 1)
        pre1 = e @ prm["fc1.weight"].T + prm["fc1.bias"]
        a = _drop(torch.relu(pre1), masks, "fc1", p)
        pre2 = a @ prm["fc2.weight"].T + prm["fc2.bias"]
        return min(float(pre1.abs().min()), float(pre2.abs().min()))


def loss_and_grads(state, x, y, masks=None, p=0.2):
    """state: dict name -> ndarray.  Returns (logits, loss, dict name -> gradient ndarray), float64.

    The loss is the reference's: F.cross_entropy(logits.transpose(1, 2), y)   (train.py:49-52)."""
    params = {k: torch.tensor(np.asarray(state[k]), dtype=torch.float64, requires_grad=True) for k in STATE_KEYS}
    logits = forward(params, x, masks, p)
    loss = F.cross_entropy(logits.transpose(1, 2), torch.as_tensor(np.asarray(y)).long())
    loss.backward()
    return (logits.detach().numpy(), float(loss.detach()),
            {k: params[k].grad.numpy() for k in STATE_KEYS})


def sample_index(n, count=512):
    """The strided positions at which the golden fixture stores a large gradient tensor."""
    if n <= count:
        return np.arange(n)
    return np.unique(np.linspace(0, n - 1, count).astype(np.int64))
