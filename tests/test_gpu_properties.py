"""Edge cases and size-independent properties of the E-/M-step kernels (GPU)."""
import copy

import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings, strategies as st

from helpers import rel
from oracle import pgpfa_oracle as po

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def make(seed, q, N, T, R, dOffset=-1.0, binSize=10):
    ex = po.synthetic_experiment(seed, q, N, R, T, binSize=binSize, dOffset=dOffset)
    ys = [t['Y'] for t in ex.data]
    rng = np.random.RandomState(seed + 1)
    params = {'C': ex.params['C'] + 0.05 * rng.randn(N, q), 'd': ex.params['d'] + 0.05 * rng.randn(N),
              'tau': ex.params['tau'] * (1 + 0.1 * rng.rand(q))}
    return ys, params


def solve(ys, params, T, **kw):
    from poisson_gpfa_b200 import core
    trials = core.DeviceTrials(dev(np.stack(ys)), 10)
    p = core.DeviceParams(params['C'], params['d'], params['tau'], T, 10)
    return trials, p, trials.estep_laplace(p, **kw)


@pytest.mark.parametrize("q,N,T,R,dOffset", [
    (1, 4, 20, 2, 0.0),       # single latent, system smaller than one tile
    (1, 5, 64, 2, 0.0),       # exactly one tile
    (1, 5, 65, 3, 0.0),       # one row spills into a second tile
    (12, 6, 11, 2, 0.0),      # maximum compiled latent dimension, N < q, odd T
    (3, 10, 129, 1, 0.0),     # a single trial, odd T (the reference's tau cost needs even T; ours does not)
    (2, 15, 40, 4, 2.5),      # high rates (tens of spikes per bin): cold-start Newton needs backtracking
])
def test_laplace_edge_shapes(q, N, T, R, dOffset):
    ys, params = make(100 + q + T, q, N, T, R, dOffset)
    trials, p, est = solve(ys, params, T)
    ir, lik, _, _ = po.laplace_struct(ys, params, T, 10, want_cov=False)
    assert rel(est.x, np.stack(ir['post_mean'])) <= 1e-8
    assert rel(est.vsm, np.stack(ir['post_vsm'])) <= 1e-8
    assert rel(est.vsmGP, np.stack([v.transpose(2, 0, 1) for v in ir['post_vsmGP']])) <= 1e-8
    assert abs(trials.post_lik(est) - lik) <= 1e-10 * abs(lik)
    # M-step on top (odd T included)
    C, d, cost, it, _ = trials.mstep_cd(p, est)
    C_o, d_o, cost_o = po.learn_Cd_newton(params, ys, ir['post_mean'], ir['post_vsm'])
    assert rel(C, C_o) <= 1e-8 and rel(d, d_o) <= 1e-8
    tau, det = trials.mstep_tau(p, trials.pautosum(est))
    tau_o, _ = po.learn_tau(params, ir, 10, gtol=1e-11)
    assert rel(tau, tau_o) <= 1e-8


def test_unsupported_latent_dimension_and_bad_inputs_fail_loudly():
    from poisson_gpfa_b200 import kernels as kn, _lib
    ys, params = make(5, 13, 4, 10, 1)
    with pytest.raises(_lib.PgpfaError):
        solve(ys, params, 10)
    y = dev(np.stack(make(6, 2, 4, 10, 2)[0]))
    with pytest.raises(AssertionError):
        kn.laplace_solve(y.transpose(1, 2), dev(np.zeros((4, 2))), dev(np.zeros(4)), dev(np.zeros((2, 10, 10))))
    with pytest.raises(AssertionError):
        kn.make_K(torch.zeros(2, dtype=torch.float64), 10, 10.0)        # host tensor: no silent CPU path


@settings(max_examples=12, deadline=None, suppress_health_check=list(HealthCheck))
@given(q=st.integers(1, 5), N=st.integers(1, 12), T=st.integers(2, 70), R=st.integers(1, 4), seed=st.integers(0, 10 ** 6))
def test_laplace_random_shapes(q, N, T, R, seed):
    ys, params = make(seed, q, N, T, R, dOffset=0.0)
    trials, p, est = solve(ys, params, T)
    ir, lik, _, _ = po.laplace_struct(ys, params, T, 10, want_cov=False)
    assert rel(est.x, np.stack(ir['post_mean'])) <= 1e-8
    assert rel(est.vsm, np.stack(ir['post_vsm'])) <= 1e-8
    assert rel(est.vsmGP, np.stack([v.transpose(2, 0, 1) for v in ir['post_vsmGP']])) <= 1e-8


def test_full_shape_properties():
    """q=8, N=100, T=200 (the headline shape) on 96 trials: properties that do not need the CPU oracle."""
    from poisson_gpfa_b200 import core, kernels as kn, _lib
    q, N, T, R = 8, 100, 200, 96
    ys, params = make(9, q, N, T, R)
    Y = np.stack(ys)
    trials, p, est = solve(ys, params, T)
    n = q * T
    # (1) stationarity: the Newton step at the returned mode is at rounding level
    f, g, W = kn.laplace_eval(est.x, trials.y, p.C, p.d, p.Kinv)
    L, D, ZT, info = kn.potrf_posterior(p.Kinv, W)
    step = kn.potrs(L, D, g.reshape(R, n))
    assert float(step.abs().max()) <= 1e-10
    assert torch.allclose(f, est.f, rtol=1e-12, atol=0)
    # (2) the slices are slices of the inverse: H * Sigma = I on a few trials, slices equal the dense inverse's
    kn.trtri(L, D, ZT, n)
    H = kn.hessian_dense(p.Kinv, W[:3].contiguous())
    Sig = kn.potri_dense(ZT[:3].contiguous(), n)
    eye = torch.eye(n, dtype=torch.float64, device="cuda")
    assert float((H @ Sig - eye).abs().max()) <= 1e-9
    idx = torch.arange(T, device="cuda")
    for r in range(3):
        for k in range(q):
            assert rel(est.vsmGP[r, k], Sig[r, k * T:(k + 1) * T, k * T:(k + 1) * T]) <= 1e-9
        blk = Sig[r].reshape(q, T, q, T)[:, idx, :, idx]          # (T, q, q)
        assert rel(est.vsm[r], blk) <= 1e-9
    # (3) trial permutation permutes the outputs bit for bit; chunked == unchunked bit for bit
    perm = np.random.RandomState(0).permutation(R)
    tr2 = core.DeviceTrials(dev(Y[perm]), 10)
    est2 = tr2.estep_laplace(p)
    pt = torch.as_tensor(perm, device="cuda")
    assert torch.equal(est2.x, est.x[pt]) and torch.equal(est2.vsm, est.vsm[pt]) and torch.equal(est2.vsmGP, est.vsmGP[pt])
    small = _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, 40)
    res3 = kn.laplace_solve(trials.y, p.C, p.d, p.Kinv, max_ws_bytes=small, lowrank=p.lowrank)
    assert res3.stats["lowrank_r"] == est.stats["lowrank_r"] > 0
    assert res3.stats["chunk"] == 40
    assert torch.equal(res3.x, est.x) and torch.equal(res3.vsm, est.vsm) and torch.equal(res3.vsmGP, est.vsmGP)
    # (4) sufficient statistics are additive over trial blocks (what the multi-GPU all-reduce relies on)
    P_all = kn.pautosum(est.vsmGP, est.x)
    P_sum = kn.pautosum(est.vsmGP[:40].contiguous(), est.x[:40].contiguous()) + \
        kn.pautosum(est.vsmGP[40:].contiguous(), est.x[40:].contiguous())
    assert rel(P_sum, P_all) <= 1e-13
    th = p.theta
    s_all = kn.mstep_cd_stats(trials.y, est.x, est.vsm, th)
    s_sum = kn.mstep_cd_stats(trials.y[:40].contiguous(), est.x[:40].contiguous(), est.vsm[:40].contiguous(), th) + \
        kn.mstep_cd_stats(trials.y[40:].contiguous(), est.x[40:].contiguous(), est.vsm[40:].contiguous(), th)
    assert rel(s_sum, s_all) <= 1e-12
    # (5) determinism: the same call twice gives identical bits
    est4 = trials.estep_laplace(p)
    assert torch.equal(est4.x, est.x) and torch.equal(est4.vsmGP, est.vsmGP)


def test_device_sampling_statistics():
    """Device-side dataset sampling (funs/util.py:733-752): not stream-compatible with numpy, so check the
    distributions: latent covariance ~ K(tau), counts ~ Poisson(exp(CX+d)) (mean and Fano factor)."""
    from poisson_gpfa_b200 import kernels as kn, util
    q, N, T, R = 3, 8, 40, 4000
    ex = util.simulate_on_device(11, q, N, R, T, binSize=10, dOffset=0.0)
    X = np.stack([t['X'] for t in ex.data])
    Y = ex.Y_all
    assert Y.shape == (R, N, T) and np.all(Y >= 0) and np.all(Y == np.round(Y))
    K = po.make_K(ex.params['tau'], T, 10)
    for k in range(q):
        emp = np.einsum('rs,rt->st', X[:, k], X[:, k]) / R
        assert np.abs(emp - K[k]).max() <= 6.0 / np.sqrt(R)                 # entries have sd <= sqrt(2/R)
    assert abs(X.mean()) <= 5.0 / np.sqrt(R * q * T) * 3
    rate = np.exp(np.einsum('nk,rkt->rnt', ex.params['C'], X) + ex.params['d'][None, :, None])
    resid = (Y - rate)
    zscore = resid.sum() / np.sqrt(rate.sum())
    assert abs(zscore) <= 5.0
    fano = (resid ** 2).sum() / rate.sum()                                   # Poisson: E (y-lam)^2 = lam
    assert abs(fano - 1.0) <= 0.02
    # reproducible for a fixed seed, different for another
    ex2 = util.simulate_on_device(11, q, N, 50, T, binSize=10, dOffset=0.0)
    ex3 = util.simulate_on_device(12, q, N, 50, T, binSize=10, dOffset=0.0)
    assert np.array_equal(ex2.Y_all, util.simulate_on_device(11, q, N, 50, T, binSize=10, dOffset=0.0).Y_all)
    assert not np.array_equal(ex2.Y_all, ex3.Y_all)
