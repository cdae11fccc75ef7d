"""CPU emulation of the operand-split schemes of the tensor-core kernels (test infrastructure: uses the oracle).

Evaluates the 3-layer bi-GRU + head of roko/rnn_model.py:57-59 with every matrix product replaced by its three-term
split -- tf32 (round 1) or fp16 with optional power-of-two operand scales (round 2, roko_b200/csrc/tc.cuh) -- and prints the
max logit error against the float64 oracle.  Run before the fp16 kernels were written; output on this box:
    fp32 1.4e-07 | tf32x3 7.0e-08 | f16x3 1.6e-07 | f16x3 scaled w*256 7.3e-08 | f16x3 scaled x16 w256 7.5e-08 (0 label mismatches)
"""
import numpy as np, torch, sys, time
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import roko_oracle as O
from roko_b200.synth import structured_windows
sd = torch.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden', 'rand_seed1.pth'), map_location='cpu')
W = {k: v.numpy() for k,v in sd.items()}
x = structured_windows(24, seed=131)
ref = O.forward(x, W, np.float64)
u64 = O.front_end(x, W, np.float64)
print('u range', u64.min(), u64.max(), 'mean', u64.mean())

def tf32(a):
    b = a.astype(np.float32).view(np.uint32).astype(np.uint64)
    b = ((b + 0x1000) & 0xFFFFE000).astype(np.uint32)   # rna to 10-bit mantissa
    return b.view(np.float32)
def split_tf32(a):
    hi = tf32(a); lo = tf32((a.astype(np.float32) - hi))
    return hi.astype(np.float64), lo.astype(np.float64)
def split_f16(a, scale=1.0):
    a = a.astype(np.float32)*np.float32(scale)
    hi = a.astype(np.float16); lo = (a - hi.astype(np.float32)).astype(np.float16)
    return hi.astype(np.float64)/scale, lo.astype(np.float64)/scale
def mm(X, Wm, split, sx=1.0, sw=1.0):
    # X (rows,k) @ Wm.T (n,k) with 3-term split; fp32 accumulate emulated by float32 cast of the final sum (products exact)
    if split is None:
        return (X.astype(np.float32) @ Wm.astype(np.float32).T).astype(np.float64)
    if split is split_f16:
        xh, xl = split(X, sx); wh, wl = split(Wm, sw)
    else:
        xh, xl = split(X); wh, wl = split(Wm)
    return (xl @ wh.T + xh @ wl.T + xh @ wh.T).astype(np.float32).astype(np.float64)

def sig(v): return 1/(1+np.exp(-v))
def gru_dir(v, Wih, Whh, bih, bhh, rev, split, sx, sw):
    B,T,_ = v.shape; H=128
    gi = (mm(v.reshape(B*T,-1), Wih, split, sx, sw).reshape(B,T,-1) + bih).astype(np.float32).astype(np.float64)
    h = np.zeros((B,H)); out = np.empty((B,T,H))
    for t in (range(T-1,-1,-1) if rev else range(T)):
        gh = mm(h, Whh, split, 1.0, sw)
        r = sig(gi[:,t,:H]+gh[:,:H]+bhh[:H]); z = sig(gi[:,t,H:2*H]+gh[:,H:2*H]+bhh[H:2*H])
        n = np.tanh(gi[:,t,2*H:] + r*(gh[:,2*H:]+bhh[2*H:]))
        h = ((1-z)*n + z*h).astype(np.float32).astype(np.float64)
        out[:,t]=h
    return out
def run(split, sx=1.0, sw=1.0):
    v = u64.astype(np.float32).astype(np.float64)
    for l in range(3):
        hs=[]
        for sfx,rev in (("",False),("_reverse",True)):
            hs.append(gru_dir(v, W[f'gru.weight_ih_l{l}{sfx}'].astype(np.float64), W[f'gru.weight_hh_l{l}{sfx}'].astype(np.float64),
                              W[f'gru.bias_ih_l{l}{sfx}'].astype(np.float64), W[f'gru.bias_hh_l{l}{sfx}'].astype(np.float64), rev, split, sx if l==0 else 1.0, sw))
        v = np.concatenate(hs,2)
    lg = v @ W['fc4.weight'].astype(np.float64).T + W['fc4.bias']
    return lg
for name, sp, sx, sw in (('fp32', None,1,1), ('tf32x3', split_tf32,1,1), ('f16x3', split_f16,1,1), ('f16x3 scaled w*256', split_f16, 1.0, 256.0), ('f16x3 scaled x16 w256', split_f16, 16.0, 256.0)):
    t=time.time(); lg = run(sp, sx, sw)
    print(f'{name:28s} max|dlogit|={np.abs(lg-ref).max():.3e}  label mismatches={(lg.argmax(2)!=ref.argmax(2)).sum()}  ({time.time()-t:.1f}s)')
print('w ranges', {k:(float(np.abs(v).max())) for k,v in W.items() if 'weight' in k})
