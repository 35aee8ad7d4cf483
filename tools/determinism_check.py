"""Bitwise run-to-run check of the three hot-path kernels (diagnostic): python tools/determinism_check.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tnmf_b200 import B200_Backend

def run(N, C, M, D, A, reps=6, **kw):
    rng = np.random.default_rng(1)
    V = rng.random((N, C) + D).astype(np.float32)
    be = B200_Backend(init='device', **kw)
    W, H = be.initialize(V, A, M, None, (-2, -1))
    H0 = H.clone()
    outs = {'recon': [], 'gradw': [], 'hupd': []}
    for i in range(reps):
        outs['recon'].append(be.reconstruct(W, H0).clone())
        g = torch.empty((2, *W.shape), dtype=W.dtype, device=W.device)
        outs['gradw'].append(be.gradient_W(V, W, H0, slice(None), g).clone())
        H.copy_(H0)
        be.update_H(V, W, H)
        outs['hupd'].append(H.clone())
        if i % 2:   # perturb timing
            torch.cuda.synchronize()
    print((N, C, M, D, A), kw, be.kernel_families(), {k: all(torch.equal(v[0], x) for x in v[1:]) for k, v in outs.items()},
          {k: max(float((v[0] - x).abs().max()) for x in v[1:]) for k, v in outs.items()})

run(6, 3, 16, (40, 56), (7, 7))
run(6, 3, 16, (40, 56), (7, 7), tmem_operand=False)
run(8, 3, 16, (256, 256), (11, 11))
run(8, 3, 16, (256, 256), (11, 11), tmem_operand=False)
