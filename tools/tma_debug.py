"""Smallest possible exercise of the three TMA kernels (debug aid; run under compute-sanitizer on the GPU box)."""
import sys
import numpy as np
import torch
sys.path.insert(0, '.')
from oracle import tnmf_oracle as orc
from tnmf_b200 import B200_Backend

which = sys.argv[1] if len(sys.argv) > 1 else 'all'
import os
N, C, M, D, A = eval(os.environ.get('TNMF_CASE', "2, 3, 4, (40, 72), (11, 11)"))
rng = np.random.default_rng(0)
V = rng.random((N, C) + D).astype(np.float32)
W = rng.random((M, C) + A).astype(np.float32)
H = rng.random((N, M) + orc.transform_shape('valid', D, A)).astype(np.float32)
be = B200_Backend(kernel_path='tma')
Wd, Hd = be.initialize(V, A, M, None, (-2, -1))
Wd.copy_(torch.from_numpy(W)); Hd.copy_(torch.from_numpy(H))
torch.cuda.synchronize()
if which in ('recon', 'all'):
    R = be.reconstruct(Wd, Hd)
    torch.cuda.synchronize()
    ref = orc.reconstruct(W.astype(np.float64), H.astype(np.float64), 'valid')
    print('recon err', np.abs(R.cpu().numpy() - ref).max() / np.abs(ref).max(), be.kernel_families())
if which in ('hupd', 'all'):
    neg, pos = be.reconstruction_gradient_H(V, Wd, Hd)
    torch.cuda.synchronize()
    rn, rp = orc.reconstruction_gradient_H(V.astype(np.float64), W.astype(np.float64), H.astype(np.float64), 'valid')
    print('hupd err', np.abs(neg.cpu().numpy() - rn).max() / np.abs(rn).max(), np.abs(pos.cpu().numpy() - rp).max() / np.abs(rp).max())
if which in ('gradw', 'all'):
    neg, pos = be.reconstruction_gradient_W(V, Wd, Hd)
    torch.cuda.synchronize()
    rn, rp = orc.reconstruction_gradient_W(V.astype(np.float64), W.astype(np.float64), H.astype(np.float64), 'valid')
    print('gradw err', np.abs(neg.cpu().numpy() - rn).max() / np.abs(rn).max(), np.abs(pos.cpu().numpy() - rp).max() / np.abs(rp).max())
