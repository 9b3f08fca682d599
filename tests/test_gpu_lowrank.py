"""GPU tests of the low-rank posterior pass (csrc/lowrank.cu): pivoted Cholesky of the smooth part of the prior, and
Sigma = eps P + Y Y^T against the dense tiled path and against the numpy oracle's dense inverse."""
import numpy as np
import pytest
import torch

from oracle import pgpfa_oracle as po
from test_gpu_kernels import dev, problem, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("T,taus", [(40, [0.05, 0.1, 0.3]), (200, list(np.linspace(0.05, 0.3, 8))), (50, [0.01, 0.6]),
                                    (300, [0.2])])
def test_prior_lowrank_factor(T, taus):
    """F F^T reproduces K - eps I to the residual tolerance; Ft is the transpose; padding columns are zero; ranks follow
    the spectrum of the squared-exponential kernel (short timescale -> full rank, long -> small)."""
    from poisson_gpfa_b200 import kernels as kn
    tau = np.asarray(taus, dtype=np.float64)
    q = len(tau)
    K = kn.make_K(dev(tau), T, 10.0, 0.001)
    F, Ft, ranks = kn.prior_lowrank(K, 0.001, 1e-14)
    Kh, Fh, Fth = K.cpu().numpy(), F.cpu().numpy(), Ft.cpu().numpy()
    assert np.array_equal(Fth, Fh.transpose(0, 2, 1))
    for k in range(q):
        assert 0 < ranks[k] <= T
        assert not Fh[k][:, ranks[k]:].any()
        S = Kh[k] - 0.001 * np.eye(T)
        assert np.abs(Fh[k] @ Fh[k].T - S).max() <= 2e-14 * T
    order = np.argsort(tau)
    assert all(ranks[order[i]] >= ranks[order[i + 1]] for i in range(q - 1))     # longer timescale, lower rank


@pytest.mark.parametrize("q,N,T,R", [(2, 20, 50, 5), (3, 7, 40, 4), (8, 100, 200, 3), (5, 3, 33, 3), (1, 5, 30, 2),
                                       (3, 15, 60, 70), (10, 30, 250, 2), (12, 20, 64, 3), (9, 12, 70, 2)])
def test_lowrank_posterior_matches_dense_path_and_oracle(q, N, T, R):
    from poisson_gpfa_b200 import kernels as kn
    ex, ys, params = problem(31 + q, q, N, T, R)
    C, d, tau = dev(params['C']), dev(params['d']), dev(params['tau'])
    y = dev(np.stack(ys))
    K = kn.make_K(tau, T, 10.0, 0.001)
    Kinv, _, _ = kn.spd_inverse(K)
    lr = kn.prior_lowrank(K, 0.001, 1e-14) + (0.001,)
    dense = kn.laplace_solve(y, C, d, Kinv, tol=1e-8)
    low = kn.laplace_solve(y, C, d, Kinv, tol=1e-8, lowrank=lr)
    assert low.stats["lowrank_r"] == sum(lr[2]) and dense.stats["lowrank_r"] == 0
    assert int(low.info.abs().max()) == 0 and low.rc == 0
    assert rel(low.x, dense.x.cpu().numpy()) <= 1e-9
    assert rel(low.f, dense.f.cpu().numpy()) <= 1e-12
    assert rel(low.vsm, dense.vsm.cpu().numpy()) <= 1e-9
    assert rel(low.vsmGP, dense.vsmGP.cpu().numpy()) <= 1e-9
    # against the dense inverse of the oracle's Hessian at the returned mode (first trials only: n^3 on the host)
    Kinv_h = Kinv.cpu().numpy()
    for r_ in range(min(R, 2)):
        xm = low.x[r_].cpu().numpy()
        H = po.assemble_H(Kinv_h, po.nlp_W_struct(xm, params['C'], params['d']))
        vsmGP_o, vsm_o = po.slice_cov(np.linalg.inv(H), q, T)
        # the covariance is that of the point before the polishing Newton step (W is not re-evaluated after it);
        # at tol = 1e-8 the two points are <= 1e-9 apart, measured deviation of the slices <= 4e-11
        assert rel(low.vsm[r_], vsm_o) <= 1e-9
        assert rel(low.vsmGP[r_], vsmGP_o.transpose(2, 0, 1)) <= 1e-9
    g = np.stack([po.nlp_grad_struct(low.x[r_].cpu().numpy(), ys[r_], params['C'], params['d'], Kinv_h) for r_ in range(min(R, 3))])
    assert np.abs(g).max() <= 1e-7      # stationary (gradient scale ~1e2-1e3)


def test_short_timescales_fall_back_to_the_dense_path():
    from poisson_gpfa_b200 import core
    p = core.DeviceParams(np.zeros((4, 2)), np.zeros(4), np.array([0.005, 0.01]), 50, 10.0)
    assert p.lowrank is None            # K - eps I has full numerical rank: nothing to gain
    p2 = core.DeviceParams(np.zeros((4, 2)), np.zeros(4), np.array([0.2, 0.4]), 50, 10.0)
    assert p2.lowrank is not None and sum(p2.lowrank[2]) < 50


@pytest.mark.parametrize("q,N,T,R", [(2, 20, 50, 5), (3, 7, 40, 1), (8, 100, 200, 9), (3, 10, 129, 3), (10, 30, 250, 4),
                                       (3, 15, 60, 70), (1, 5, 30, 2), (12, 20, 64, 3)])
def test_pautosum_inside_the_lowrank_pass(q, N, T, R):
    """makePrecomp's PautoSum (funs/learning.py:162-165) taken inside the posterior pass as ONE symmetric product per
    latent over all trials (syrk_sum_kernel) against the sum over the per-trial post_vsmGP blocks of the same pass, and
    against the oracle's sum over dense inverses."""
    from poisson_gpfa_b200 import kernels as kn
    ex, ys, params = problem(57 + q, q, N, T, R)
    C, d, tau = dev(params['C']), dev(params['d']), dev(params['tau'])
    y = dev(np.stack(ys))
    K = kn.make_K(tau, T, 10.0, 0.001)
    Kinv, _, _ = kn.spd_inverse(K)
    lr = kn.prior_lowrank(K, 0.001, 1e-14) + (0.001,)
    both = kn.laplace_solve(y, C, d, Kinv, lowrank=lr, want_vsmGP=True, want_pautosum=True)
    only = kn.laplace_solve(y, C, d, Kinv, lowrank=lr, want_vsmGP=False, want_pautosum=True)
    assert only.vsmGP is None and both.stats["lowrank_r"] > 0
    ref = kn.pautosum(both.vsmGP, both.x)
    assert rel(both.pautosum, ref.cpu().numpy()) <= 1e-13
    assert torch.equal(only.pautosum, both.pautosum) and torch.equal(only.x, both.x)
    assert rel(both.pautosum, both.pautosum.transpose(1, 2).cpu().numpy()) == 0.0      # exactly symmetric
    ir, _, _, _ = po.laplace_struct(ys, params, T, 10, want_cov=False)
    P_o = np.stack([pp['PautoSum'] for pp in po.make_precomp(ir)])
    assert rel(both.pautosum, P_o) <= 1e-9


def test_pautosum_chunked_pass_accumulates():
    from poisson_gpfa_b200 import kernels as kn, _lib
    q, N, T, R = 3, 9, 48, 11
    ex, ys, params = problem(3, q, N, T, R)
    C, d, tau = dev(params['C']), dev(params['d']), dev(params['tau'])
    y = dev(np.stack(ys))
    K = kn.make_K(tau, T, 10.0, 0.001)
    Kinv, _, _ = kn.spd_inverse(K)
    lr = kn.prior_lowrank(K, 0.001, 1e-14) + (0.001,)
    full = kn.laplace_solve(y, C, d, Kinv, lowrank=lr, want_vsmGP=False, want_pautosum=True)
    small = _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, 4)
    chunked = kn.laplace_solve(y, C, d, Kinv, lowrank=lr, want_vsmGP=False, want_pautosum=True, max_ws_bytes=small)
    assert chunked.stats["chunk"] < R
    assert rel(chunked.pautosum, full.pautosum.cpu().numpy()) <= 1e-13
