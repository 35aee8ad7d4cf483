"""
cfg4-style streamed fit, timed end to end: 1-D signals held in PINNED HOST memory are cut into
subsamples by `fit_stream` (tnmf/TransformInvariantNMF.py:506-523), each subsample is copied to the device on a side
stream while the previous one is being fitted with Cyclic MU over minibatches (:457-465), W is kept from subsample
to subsample.  Prints one JSON line; the timed region contains every host-to-device copy.

    python tools/stream_bench.py [--n-signals 16384] [--length 4096] [--atoms 64] [--width 128]
                                 [--subsample 4096] [--batch 1024] [--epochs 3]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/stream_bench.py ...
        one process per GPU: every rank streams its own `--n-signals` signals, the W gradient is all-reduced once per
        epoch, the rate is the job's (all ranks' signals over the slowest rank's time)

`--n-signals` defaults to 16384 (268 MB of pinned host memory); BASELINE config 4 as written is `--n-signals 131072` on
8 GPUs (1 M signals, 2.1 GB of pinned host memory per rank): profiles/r02_stream_cfg4_n8_1M.json.
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
    ap.add_argument('--n-signals', dest='signals', type=int, default=16384)
    ap.add_argument('--length', type=int, default=4096)
    ap.add_argument('--atoms', type=int, default=64)
    ap.add_argument('--width', type=int, default=128)
    ap.add_argument('--subsample', type=int, default=4096)
    ap.add_argument('--batch', type=int, default=1024)
    ap.add_argument('--epochs', type=int, default=3)
    ap.add_argument('--init', default='device', choices=['device', 'numpy'],
                    help="'numpy' = the reference's float64 host draw of H per subsample (8.8 GB and ~9 s per 4096 signals)")
    a = ap.parse_args()

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    if world > 1:           # torchrun: one process per GPU, every rank streams its own `--n-signals` signals
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', '0')))
        torch.distributed.init_process_group('nccl')
    g = torch.Generator().manual_seed(rank)
    V = torch.rand((a.signals, 1, a.length), generator=g, dtype=torch.float32).pin_memory()
    kw = dict(algorithm=MiniBatchAlgorithm.Cyclic_MU, subsample_size=a.subsample, batch_size=a.batch,
              n_epochs=a.epochs, progress_callback=lambda *_: True)

    def run(source):
        np.random.seed(0)
        nmf = TransformInvariantNMF(n_atoms=a.atoms, atom_shape=(a.width,), backend='b200', init=a.init,
                                    input_is_local_shard=world > 1, equal_shards=world > 1)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        nmf.fit(source, **kw)
        w = nmf.W                                 # device -> host read of the dictionary ends the timed region
        torch.cuda.synchronize()
        return time.perf_counter() - t0, w, nmf

    run(V[:2 * a.subsample])                      # warm-up: module load, pinned staging, allocator
    dt, w, nmf = run(V)
    if world > 1:           # the job is as slow as its slowest rank
        t = torch.tensor([dt], dtype=torch.float64, device='cuda')
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        dt = float(t.item())
        w0 = torch.from_numpy(w).cuda()
        torch.distributed.broadcast(w0, 0)
        assert torch.equal(w0.cpu(), torch.from_numpy(w)), 'the dictionary differs between ranks'
    n_sub = -(-a.signals // a.subsample)
    sample_iters = a.signals * a.epochs * world
    line = {
        'metric': 'sample-iterations/sec, streamed Cyclic-MU fit from pinned host memory (cfg4 shape)',
        'value': sample_iters / dt, 'unit': 'sample-iterations/s', 'n_gpus': world, 'seconds': dt,
        'config': {'signals_per_gpu': a.signals, 'length': a.length, 'atoms': a.atoms, 'atom_width': a.width,
                   'subsample_size': a.subsample, 'batch_size': a.batch, 'n_epochs': a.epochs, 'subsamples': n_sub, 'init': a.init},
        'h2d_bytes_per_gpu': int(V.numel() * 4), 'h2d_gbs_per_gpu': V.numel() * 4 / dt / 1e9,
        'kernel_path': nmf._backend.kernel_families(),
        'W_finite': bool(np.isfinite(w).all()), 'W_rows_sum_to_one': bool(np.allclose(w.sum(axis=-1), 1, atol=1e-4)),
    }
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:           # leave without tearing NCCL down under the captured step graphs (see bench.py finish_ranks)
        torch.cuda.synchronize()
        torch.distributed.barrier()
        sys.stdout.flush()
        os._exit(0)


if __name__ == '__main__':
    main()
