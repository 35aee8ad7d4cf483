"""
cfg4-style streamed fit, timed end to end: 1-D signals held in PINNED HOST memory are cut into
subsamples by `fit_stream` (tnmf/TransformInvariantNMF.py:506-523), each subsample is copied to the device on a side
stream while the previous one is being fitted with Cyclic MU over minibatches (:457-465), W is kept from subsample
to subsample.  Prints one JSON line; the timed region contains every host-to-device copy.

    python tools/stream_bench.py [--signals 16384] [--length 4096] [--atoms 64] [--width 128]
                                 [--subsample 4096] [--batch 1024] [--epochs 3]

`--signals` defaults to 16384 (268 MB of pinned host memory), not BASELINE's 1 M (16 GB): the rate is per signal and
the stream is consumed subsample by subsample, so the total length only changes the run time.
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tnmf_b200 import MiniBatchAlgorithm, TransformInvariantNMF      # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--signals', type=int, default=16384)
    ap.add_argument('--length', type=int, default=4096)
    ap.add_argument('--atoms', type=int, default=64)
    ap.add_argument('--width', type=int, default=128)
    ap.add_argument('--subsample', type=int, default=4096)
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--epochs', type=int, default=3)
    ap.add_argument('--init', default='device', choices=['device', 'numpy'],
                    help="'numpy' = the reference's float64 host draw of H per subsample (8.8 GB and ~9 s per 4096 signals)")
    a = ap.parse_args()

    g = torch.Generator().manual_seed(0)
    V = torch.rand((a.signals, 1, a.length), generator=g, dtype=torch.float32).pin_memory()
    kw = dict(algorithm=MiniBatchAlgorithm.Cyclic_MU, subsample_size=a.subsample, batch_size=a.batch,
              n_epochs=a.epochs, progress_callback=lambda *_: True)

    def run(source):
        np.random.seed(0)
        nmf = TransformInvariantNMF(n_atoms=a.atoms, atom_shape=(a.width,), backend='b200', init=a.init)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nmf.fit(source, **kw)
        w = nmf.W                                 # device -> host read of the dictionary ends the timed region
        torch.cuda.synchronize()
        return time.perf_counter() - t0, w, nmf

    run(V[:2 * a.subsample])                      # warm-up: module load, pinned staging, allocator
    dt, w, nmf = run(V)
    n_sub = -(-a.signals // a.subsample)
    sample_iters = a.signals * a.epochs
    line = {
        'metric': 'sample-iterations/sec, streamed Cyclic-MU fit from pinned host memory (cfg4 shape)',
        'value': sample_iters / dt, 'unit': 'sample-iterations/s', 'n_gpus': 1, 'seconds': dt,
        'config': {'signals': a.signals, 'length': a.length, 'atoms': a.atoms, 'atom_width': a.width,
                   'subsample_size': a.subsample, 'batch_size': a.batch, 'n_epochs': a.epochs, 'subsamples': n_sub, 'init': a.init},
        'h2d_bytes': int(V.numel() * 4), 'h2d_gbs_needed': V.numel() * 4 / dt / 1e9,
        'kernel_path': nmf._backend.kernel_families(),
        'W_finite': bool(np.isfinite(w).all()), 'W_rows_sum_to_one': bool(np.allclose(w.sum(axis=-1), 1, atol=1e-4)),
    }
    print(json.dumps(line))


if __name__ == '__main__':
    main()
