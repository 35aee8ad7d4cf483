import sys, time, torch, numpy as np
sys.path.insert(0, '.')
from tnmf_b200 import TransformInvariantNMF
def run(shape, atom, its=10, **kw):
    V = torch.rand(shape, device='cuda', dtype=torch.float32)
    nmf = TransformInvariantNMF(n_atoms=64, atom_shape=atom, backend='b200', init='device', **kw)
    nmf.fit_batch(V, n_iterations=3, progress_callback=lambda *_: True)
    torch.cuda.synchronize(); t0 = time.perf_counter()
    nmf.fit_batch(V, n_iterations=its, progress_callback=lambda *_: True)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(shape, atom, kw, nmf._backend.kernel_families(), 'ms/iter (incl. init)', dt / its * 1e3, flush=True)
run((2048, 1, 4096), (128,))
run((2048, 1, 4096), (128,), its=30)
for path in ('auto', 'tma', 'tiled'):
    try:
        run((1, 1, 2048, 4096), (1, 128), kernel_path=path)
        run((1, 1, 2048, 4096), (1, 128), its=30, kernel_path=path)
    except Exception as e:
        print(path, 'failed', repr(e)[:300], flush=True)
