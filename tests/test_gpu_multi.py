"""Two GPUs, two processes (NCCL): the trial-sharded EM iteration through the CUDA kernels must equal the single-GPU
one ("multi-GPU = single-GPU bitwise-or-1e-12").  Skipped when fewer than two devices are visible; run
with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` (record in profiles/)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _problem():
    sys.path.insert(0, ROOT)
    from poisson_gpfa_b200 import util
    q, N, T, R = 3, 15, 60, 9                       # 9 trials over 2 ranks: shards of 5 and 4
    ex = util.simulate(11, q, N, R, T, binSize=10, dOffset=0.0)
    Y = np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data])
    rng = np.random.RandomState(3)
    ip = {'C': ex.params['C'] + 0.05 * rng.randn(N, q), 'd': ex.params['d'] + 0.05 * rng.randn(N), 'tau': ex.params['tau'] * 1.1}
    return Y, ip, T


def _run(trials, ip, T, n_iter=3):
    from poisson_gpfa_b200 import core
    params = core.DeviceParams(ip['C'], ip['d'], ip['tau'], T, 10)
    x0, out = None, []
    for _ in range(n_iter):
        params, est, lik, info = trials.em_step(params, x0=x0)
        x0 = est.x
        out.append((lik, params.C.cpu().numpy(), params.d.cpu().numpy(), params.tau.cpu().numpy(), info["cd_iters"]))
    return out, est


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    from poisson_gpfa_b200 import core, dist, _lib
    red = dist.init_from_env()
    assert red.world_size == world and torch.distributed.get_backend() == "nccl"
    Y, ip, T = _problem()
    lo, hi = dist.shard_bounds(Y.shape[0], world, rank)
    trials = core.DeviceTrials(_lib.dev_f64(Y[lo:hi]), 10, red, R_total=Y.shape[0], offset=lo)
    out, est = _run(trials, ip, T)
    # an empty shard on one rank (mini-batch smaller than the world) must join the collectives with zeros
    sel = slice(0, 1) if rank == 0 else slice(0, 0)
    t1 = core.DeviceTrials(_lib.dev_f64(Y[sel]), 10, red, R_total=1, offset=0)
    out1, _ = _run(t1, ip, T, n_iter=1)
    if rank == 0:
        np.savez(out_path, lik=[o[0] for o in out], C=np.stack([o[1] for o in out]), d=np.stack([o[2] for o in out]),
                 tau=np.stack([o[3] for o in out]), x=est.x.cpu().numpy(), n_allreduce=red.n_allreduce,
                 lik1=out1[0][0], C1=out1[0][1], tau1=out1[0][3])
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()


@pytest.mark.timeout(600)
def test_two_gpu_em_equals_single_gpu(tmp_path):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    out = str(tmp_path / "res.npz")
    mp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    res = np.load(out)
    sys.path.insert(0, ROOT)
    from poisson_gpfa_b200 import core, _lib
    Y, ip, T = _problem()
    single, est = _run(core.DeviceTrials(_lib.dev_f64(Y), 10), ip, T)
    rel = lambda a, b: float(np.abs(np.asarray(a) - np.asarray(b)).max() / np.abs(np.asarray(b)).max())
    for it, (lik, C, d, tau, _) in enumerate(single):
        assert abs(res['lik'][it] - lik) <= 1e-12 * abs(lik)
        assert rel(res['C'][it], C) <= 1e-12 and rel(res['d'][it], d) <= 1e-12 and rel(res['tau'][it], tau) <= 1e-12
    assert rel(res['x'], est.x[:res['x'].shape[0]].cpu().numpy()) <= 1e-12
    one, _ = _run(core.DeviceTrials(_lib.dev_f64(Y[:1]), 10), ip, T, n_iter=1)
    assert abs(res['lik1'] - one[0][0]) <= 1e-12 * abs(one[0][0])
    assert rel(res['C1'], one[0][1]) <= 1e-12 and rel(res['tau1'], one[0][3]) <= 1e-12
