"""N>1 host logic on CPU: two gloo ranks shard the trials, reduce the sufficient statistics through
poisson_gpfa_b200.dist.Reducer, and must reproduce the unsharded result (the per-shard statistics come
from the oracle here; on GPUs they come from the CUDA kernels, tests/test_gpu_em.py)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_path):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    import torch.distributed as td
    from poisson_gpfa_b200 import dist
    from oracle import pgpfa_oracle as po
    red = dist.init_from_env()
    assert red.world_size == world and red.rank == rank and td.get_backend() == "gloo"
    q, N, T, R = 2, 6, 20, 7
    ex = po.synthetic_experiment(5, q, N, R, T, dOffset=0.0)
    ys = [t['Y'] for t in ex.data]
    params = {'C': ex.params['C'] * 0.9, 'd': ex.params['d'] + 0.1, 'tau': ex.params['tau'] * 1.1}
    lo, hi = dist.shard_bounds(R, world, rank)
    ir, lik, _, _ = po.laplace_struct(ys[lo:hi], params, T, 10, want_cov=False)
    # (a) objective sum, (b) PautoSum, (c) per-neuron Newton statistics  
    f_sum = red.sum_scalar(-lik * (hi - lo))
    pre = po.make_precomp(ir)
    P = red.sum_tensor(torch.from_numpy(np.stack([p['PautoSum'] for p in pre])))
    f, g, H = po.obs_stats_struct(params['C'], params['d'], ys[lo:hi], ir['post_mean'], ir['post_vsm'])
    stats = red.sum_tensor(torch.from_numpy(np.concatenate([f[:, None], g, H.reshape(N, -1)], axis=1)))
    n_tot = red.sum_scalar(hi - lo)
    mx = red.max_scalar(float(rank))
    if rank == 0:
        np.savez(out_path, f_sum=f_sum, P=P.numpy(), stats=stats.numpy(), n_tot=n_tot, mx=mx, n_allreduce=red.n_allreduce)
    td.barrier()
    td.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharded_statistics_match_unsharded(tmp_path):
    from oracle import pgpfa_oracle as po
    out = str(tmp_path / "res.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    res = np.load(out)
    q, N, T, R = 2, 6, 20, 7
    ex = po.synthetic_experiment(5, q, N, R, T, dOffset=0.0)
    ys = [t['Y'] for t in ex.data]
    params = {'C': ex.params['C'] * 0.9, 'd': ex.params['d'] + 0.1, 'tau': ex.params['tau'] * 1.1}
    ir, lik, _, _ = po.laplace_struct(ys, params, T, 10, want_cov=False)
    assert abs(res['f_sum'] - (-lik * R)) <= 1e-12 * abs(lik * R)
    P = np.stack([p['PautoSum'] for p in po.make_precomp(ir)])
    assert np.abs(res['P'] - P).max() <= 1e-12 * np.abs(P).max()
    f, g, H = po.obs_stats_struct(params['C'], params['d'], ys, ir['post_mean'], ir['post_vsm'])
    full = np.concatenate([f[:, None], g, H.reshape(N, -1)], axis=1)
    assert np.abs(res['stats'] - full).max() <= 1e-12 * np.abs(full).max()
    assert res['n_tot'] == R and res['mx'] == 1.0 and res['n_allreduce'] == 4
