"""Runs the three hot-path operations of one workload once (after a warm-up) - for the -DTNMF_TC_PROFILE build, whose
kernels print where each role waited:  TNMF_LIB_PATH=tools/prof/libtnmf_prof.so python tools/prof_kernels.py [cfg2]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tnmf_b200 import B200_Backend

name = sys.argv[1] if len(sys.argv) > 1 else 'cfg2'
w = bench.WORKLOADS[name]
be = B200_Backend(init='device')
V = torch.rand((w['N'], w['C'], *w['D']), device=be.device)
W, H = be.initialize(V, w['A'], w['M'], None, tuple(range(-len(w['A']), 0)))
g = torch.empty((2, *W.shape), dtype=W.dtype, device=W.device)
for i in range(2):
    print(f'--- pass {i}', flush=True)
    be.reconstruct(W, H)
    torch.cuda.synchronize()
    be.gradient_W(V, W, H, slice(None), g)
    torch.cuda.synchronize()
    be.update_H(V, W, H)
    torch.cuda.synchronize()
print(be.kernel_families())
