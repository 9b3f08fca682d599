"""GPU parity tests of the individual kernels against the numpy oracle (same seeded inputs).
Tolerances are relative to the largest magnitude of the reference quantity and are written per test."""
import numpy as np
import pytest
import torch

from oracle import pgpfa_oracle as po

pytestmark = pytest.mark.gpu


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).cuda()


def random_spd(rng, b, n, cond=1e3):
    out = np.empty((b, n, n))
    for i in range(b):
        Qm, _ = np.linalg.qr(rng.randn(n, n))
        ev = np.exp(rng.uniform(0, np.log(cond), n))
        out[i] = (Qm * ev) @ Qm.T
        out[i] = 0.5 * (out[i] + out[i].T)
    return out


def problem(seed, q, N, T, R, binSize=10, dOffset=-1.0):
    ex = po.synthetic_experiment(seed, q, N, R, T, binSize=binSize, dOffset=dOffset)
    ys = [tr['Y'] for tr in ex.data]
    rng = np.random.RandomState(seed + 1)
    params = {'C': ex.params['C'] + 0.05 * rng.randn(N, q), 'd': ex.params['d'] + 0.05 * rng.randn(N),
              'tau': ex.params['tau'] * (1 + 0.1 * rng.rand(q))}
    return ex, ys, params


@pytest.mark.parametrize("q,T,bs", [(2, 50, 20), (3, 40, 10), (8, 200, 10), (10, 250, 10), (1, 7, 10)])
def test_make_K(q, T, bs):
    from poisson_gpfa_b200 import kernels as kn
    tau = np.linspace(0.05, 0.4, q)
    K = kn.make_K(dev(tau), T, bs)
    assert rel(K, po.make_K(tau, T, bs)) <= 1e-15
    Kb = kn.make_K_big(K)
    Kb_o, _ = po.make_K_big({'C': np.zeros((3, q)), 'tau': tau}, T * bs, bs)
    assert rel(Kb, Kb_o) <= 1e-15


@pytest.mark.parametrize("n,b", [(40, 3), (64, 2), (100, 5), (200, 8), (250, 2), (130, 1)])
def test_spd_inverse(n, b):
    from poisson_gpfa_b200 import kernels as kn
    rng = np.random.RandomState(n)
    A = random_spd(rng, b, n)
    Ainv, logdet, info = kn.spd_inverse(dev(A))
    assert int(info.abs().max()) == 0
    assert rel(Ainv, np.linalg.inv(A)) <= 1e-11       # cond 1e3
    assert rel(logdet, np.linalg.slogdet(A)[1]) <= 1e-13


@pytest.mark.parametrize("n,b", [(1, 2), (7, 3), (10, 1), (11, 4), (199, 2), (208, 72), (220, 3), (221, 2)])
def test_spd_inverse_register_resident_sweep_and_its_limits(n, b):
    """n <= 220 goes through the one-CTA-per-matrix sweep (gpprior.cu: spd_sweep_kernel; 10 x 10 register tiles, ragged
    last tile padded with the identity), n = 221 through the tile-kernel chain: same results, symmetric output."""
    from poisson_gpfa_b200 import kernels as kn
    rng = np.random.RandomState(1000 + n)
    A = random_spd(rng, b, n)
    Ainv, logdet, info = kn.spd_inverse(dev(A))
    assert int(info.abs().max()) == 0
    assert rel(Ainv, np.linalg.inv(A)) <= 1e-11
    ld = np.linalg.slogdet(A)[1]
    assert np.abs(logdet.cpu().numpy() - ld).max() <= 1e-12 * max(1.0, np.abs(ld).max())
    assert torch.equal(Ainv, Ainv.transpose(1, 2))


def test_spd_inverse_reports_the_first_non_positive_pivot():
    from poisson_gpfa_b200 import kernels as kn
    rng = np.random.RandomState(5)
    A = random_spd(rng, 3, 40)
    A[1, 7, 7] = -1.0                     # the 8th pivot of matrix 1 is negative
    _, _, info = kn.spd_inverse(dev(A))
    assert info.cpu().numpy().tolist() == [0, 8, 0]


def test_kinv_of_prior():
    from poisson_gpfa_b200 import kernels as kn
    tau = np.linspace(0.05, 0.3, 8)
    K = po.make_K(tau, 200, 10)
    Kinv, logdet, info = kn.spd_inverse(kn.make_K(dev(tau), 200, 10))
    assert int(info.abs().max()) == 0
    # cond(K) ~ 1e5: both inverses carry ~1e-11 relative rounding
    assert rel(Kinv, np.stack([np.linalg.inv(K[k]) for k in range(8)])) <= 1e-9
    assert rel(logdet, np.linalg.slogdet(K)[1]) <= 1e-12


@pytest.mark.parametrize("q,T,R", [(1, 5, 1), (3, 70, 37), (8, 200, 70), (2, 129, 33)])
def test_prior_apply_is_a_batched_matrix_product(q, T, R):
    """out[r,k,:] = Kmat[k] @ v[r,k,:] for general (also non-symmetric) matrices, ragged tile edges included."""
    from poisson_gpfa_b200 import kernels as kn
    rng = np.random.RandomState(q * 1000 + T)
    Kmat, v = rng.randn(q, T, T), rng.randn(R, q, T)
    out = kn.prior_apply(dev(Kmat), dev(v))
    assert rel(out, np.einsum('kst,rkt->rks', Kmat, v)) <= 1e-13


@pytest.mark.parametrize("n,b", [(100, 3), (192, 2), (450, 2), (1600, 2)])
def test_factor_solve_invert(n, b):
    from poisson_gpfa_b200 import kernels as kn
    rng = np.random.RandomState(n + 7)
    A = random_spd(rng, b, n)
    L, D, ZT, info = kn.potrf_dense(dev(A))
    assert int(info.abs().max()) == 0
    Ld = kn.tiles_to_dense(L, n)
    assert rel(Ld, np.linalg.cholesky(A)) <= 1e-12
    rhs = rng.randn(b, n)
    x = kn.potrs(L, D, dev(rhs), scale=-1.0)
    assert rel(x, -np.linalg.solve(A, rhs[:, :, None])[:, :, 0]) <= 1e-11
    kn.trtri(L, D, ZT, n)
    Zd = kn.tiles_to_dense(ZT, n, upper=True)
    assert rel(Zd, np.linalg.inv(np.linalg.cholesky(A)).transpose(0, 2, 1)) <= 1e-11
    Ainv = kn.potri_dense(ZT, n)
    assert rel(Ainv, np.linalg.inv(A)) <= 1e-11
    assert rel(kn.logdet(L, n), np.linalg.slogdet(A)[1]) <= 1e-13


def test_not_spd_is_reported():
    from poisson_gpfa_b200 import kernels as kn
    A = np.eye(70)[None].copy()
    A[0, 66, 66] = -1.0
    _, _, _, info = kn.potrf_dense(dev(A))
    assert int(info[0]) == 67


@pytest.mark.parametrize("q,N,T,R", [(2, 20, 50, 5), (3, 7, 40, 4), (8, 100, 200, 3), (10, 30, 250, 2), (1, 5, 30, 2),
                                       (5, 3, 33, 3)])
def test_laplace_eval_and_posterior_factor(q, N, T, R):
    from poisson_gpfa_b200 import kernels as kn
    ex, ys, params = problem(11 + q, q, N, T, R)
    C, d = params['C'], params['d']
    K = po.make_K(params['tau'], T, 10)
    Kinv = np.stack([np.linalg.inv(K[k]) for k in range(q)])
    rng = np.random.RandomState(5)
    X = 0.3 * rng.randn(R, q, T)
    f, g, W = kn.laplace_eval(dev(X), dev(np.stack(ys)), dev(C), dev(d), dev(Kinv))
    f_o = np.array([po.nlp_struct(X[r], ys[r], C, d, Kinv) for r in range(R)])
    g_o = np.stack([po.nlp_grad_struct(X[r], ys[r], C, d, Kinv) for r in range(R)])
    W_o = np.stack([po.nlp_W_struct(X[r], C, d).transpose(1, 2, 0).reshape(q * q, T) for r in range(R)])
    assert rel(f, f_o) <= 1e-13
    assert rel(g, g_o) <= 1e-11       # Kinv x with cond(K) ~ 1e5 and random x
    assert rel(W, W_o) <= 1e-13
    H_o = np.stack([po.assemble_H(Kinv, po.nlp_W_struct(X[r], C, d)) for r in range(R)])
    assert rel(kn.hessian_dense(dev(Kinv), W), H_o) <= 1e-13
    # factor H without materialising it; inverse slices vs numpy inverse of the dense oracle Hessian
    L, D, ZT, info = kn.potrf_posterior(dev(Kinv), W)
    assert int(info.abs().max()) == 0
    n = q * T
    assert rel(kn.tiles_to_dense(L, n), np.linalg.cholesky(H_o)) <= 1e-11
    kn.trtri(L, D, ZT, n)
    vsm, vsmGP = kn.cov_slices(ZT, q, T)
    cov = np.linalg.inv(H_o)
    sl = [po.slice_cov(cov[r], q, T) for r in range(R)]
    # cond(H) reaches ~1e5 when few neurons inform the posterior: both inverses carry cond*eps rounding,
    # so the bound is the north-star 1e-8 here (well-conditioned cases land at 1e-12)
    assert rel(vsm, np.stack([s[1] for s in sl])) <= 1e-8
    assert rel(vsmGP, np.stack([s[0].transpose(2, 0, 1) for s in sl])) <= 1e-8
    assert rel(kn.potri_dense(ZT, n), cov) <= 1e-8
    step = kn.potrs(L, D, g.reshape(R, n), scale=-1.0)
    assert rel(step, -np.linalg.solve(H_o, g_o.reshape(R, n, 1))[:, :, 0]) <= 1e-8
    # backward-stable check that does not depend on numpy's inverse: H * Sigma = I
    Sig = kn.potri_dense(ZT, n).cpu().numpy()
    assert np.abs(H_o @ Sig - np.eye(n)).max() <= 1e-9


@pytest.mark.parametrize("q,N,T,R", [(2, 20, 50, 5), (3, 7, 40, 4), (8, 100, 200, 4), (5, 3, 33, 3)])
def test_laplace_solve_fixed_point(q, N, T, R):
    from poisson_gpfa_b200 import kernels as kn
    ex, ys, params = problem(21 + q, q, N, T, R)
    C, d = params['C'], params['d']
    K = po.make_K(params['tau'], T, 10)
    Kinv = np.stack([np.linalg.inv(K[k]) for k in range(q)])
    res = kn.laplace_solve(dev(np.stack(ys)), dev(C), dev(d), dev(Kinv), want_cov=True)
    assert res.rc == 0 and int(res.info.abs().max()) == 0
    ir, lik, optim, iters = po.laplace_struct(ys, params, T, 10)
    # tolerance: north-star 1e-8 relative on posterior means / covariances / objective
    assert rel(res.x, np.stack(ir['post_mean'])) <= 1e-8
    assert rel(res.cov, np.stack(ir['post_cov'])) <= 1e-8
    assert rel(res.vsm, np.stack(ir['post_vsm'])) <= 1e-8
    assert rel(res.vsmGP, np.stack([v.transpose(2, 0, 1) for v in ir['post_vsmGP']])) <= 1e-8
    assert abs(-float(res.f.mean()) - lik) <= 1e-10 * abs(lik)
    # warm start at the mode converges in one Newton iteration and stays there
    res2 = kn.laplace_solve(dev(np.stack(ys)), dev(C), dev(d), dev(Kinv), x0=res.x, want_vsm=False, want_vsmGP=False)
    assert int(res2.niter.max()) == 1
    assert rel(res2.x, np.stack(ir['post_mean'])) <= 1e-8


def test_laplace_solve_chunked_equals_unchunked():
    from poisson_gpfa_b200 import kernels as kn, _lib
    q, N, T, R = 3, 9, 40, 7
    ex, ys, params = problem(3, q, N, T, R)
    K = po.make_K(params['tau'], T, 10)
    Kinv = dev(np.stack([np.linalg.inv(K[k]) for k in range(q)]))
    y, C, d = dev(np.stack(ys)), dev(params['C']), dev(params['d'])
    a = kn.laplace_solve(y, C, d, Kinv)
    small = _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, 3)
    b = kn.laplace_solve(y, C, d, Kinv, max_ws_bytes=small)
    assert b.stats["chunk"] == 3
    assert torch.equal(a.x, b.x) and torch.equal(a.vsm, b.vsm) and torch.equal(a.vsmGP, b.vsmGP)


@pytest.mark.parametrize("q,N,T,R", [(2, 20, 50, 5), (3, 7, 40, 4), (8, 100, 200, 3), (10, 200, 250, 2), (4, 300, 20, 2)])
def test_mstep_cd_stats(q, N, T, R):
    from poisson_gpfa_b200 import kernels as kn
    ex, ys, params = problem(31 + q, q, N, T, R)
    rng = np.random.RandomState(1)
    means = [0.5 * rng.randn(q, T) for _ in range(R)]
    vsms = []
    for _ in range(R):
        A = 0.2 * rng.randn(T, q, q)
        vsms.append(A @ A.transpose(0, 2, 1) + 0.01 * np.eye(q))
    C, d = params['C'], params['d']
    theta = np.concatenate([C, d[:, None]], axis=1)
    st = kn.mstep_cd_stats(dev(np.stack(ys)), dev(np.stack(means)), dev(np.stack(vsms)), dev(theta)).cpu().numpy()
    f, g, H = po.obs_stats_struct(C, d, ys, means, vsms)
    assert rel(st[0], f) <= 1e-13
    assert rel(st[1:q + 2].T, g) <= 1e-12
    iu = np.triu_indices(q + 1)
    assert rel(st[q + 2:].T, H[:, iu[0], iu[1]]) <= 1e-12


@pytest.mark.parametrize("q,T,R", [(2, 50, 5), (3, 40, 4), (8, 200, 6)])
def test_pautosum_and_tau_eval(q, T, R):
    from poisson_gpfa_b200 import kernels as kn
    rng = np.random.RandomState(q)
    means = rng.randn(R, q, T)
    tau = np.linspace(0.05, 0.3, q)
    K = po.make_K(tau, T, 10)
    vs = np.stack([np.stack([0.3 * K[k] + 0.01 * np.eye(T) for k in range(q)]) for _ in range(R)])
    P = kn.pautosum(dev(vs), dev(means))
    infRes = {'post_mean': list(means), 'post_vsmGP': [v.transpose(1, 2, 0) for v in vs]}
    pre = po.make_precomp(infRes)
    assert rel(P, np.stack([p['PautoSum'] for p in pre])) <= 1e-14
    p = np.log(1.0 / (tau * 1000 / 10) ** 2) + 0.2 * rng.randn(q)
    cost, grad = kn.tau_eval(dev(p), P, R, T)
    c_o = np.array([po.tau_cost(p[k], pre[k]) for k in range(q)])
    g_o = np.array([po.tau_cost_grad(p[k], pre[k])[0] for k in range(q)])
    assert rel(cost, c_o) <= 1e-11
    assert rel(grad, g_o) <= 1e-8        # gradient is a difference of two large traces (cond(K) ~ 1e5)
    cost2, grad2 = kn.tau_eval(dev(p), P, R, T, prior_w=1 / 0.3 ** 2, tau_old=dev(tau), binSize=10)
    c2 = np.array([po.tau_cost_prior(p[k], pre[k], 10, tau[k], 0.3) for k in range(q)])
    g2 = np.array([po.tau_cost_prior_grad(p[k], pre[k], 10, tau[k], 0.3)[0] for k in range(q)])
    assert rel(cost2, c2) <= 1e-11
    assert rel(grad2, g2) <= 1e-8


def test_inexact_newton_pcg_reaches_same_fixed_point():
    """The inexact-Newton path (conjugate gradients with the block-Jacobi preconditioner, no qT x qT factorisation
    before the mode) and the exact-Newton path (fresh factorisation per iteration) must land on the same posterior."""
    from poisson_gpfa_b200 import core
    q, N, T, R = 4, 30, 100, 6
    ex, ys, params = problem(77, q, N, T, R)
    trials = core.DeviceTrials(dev(np.stack(ys)), 10)
    pA = core.DeviceParams(params['C'], params['d'], params['tau'], T, 10)
    ir, lik, _, _ = po.laplace_struct(ys, params, T, 10, want_cov=False)
    cold_pcg = trials.estep_laplace(pA)
    cold_newton = trials.estep_laplace(pA, inexact_newton=False)
    assert cold_pcg.stats["pcg_iters"] > 0 and cold_pcg.stats["factorizations"] == R       # only the one at the mode
    assert cold_newton.stats["pcg_iters"] == 0 and cold_newton.stats["factorizations"] > R
    for est in (cold_pcg, cold_newton):
        assert rel(est.x, np.stack(ir['post_mean'])) <= 1e-8
        assert rel(est.vsm, np.stack(ir['post_vsm'])) <= 1e-8
        assert rel(est.vsmGP, np.stack([v.transpose(2, 0, 1) for v in ir['post_vsmGP']])) <= 1e-8
    # warm start under nearby and under far-away parameters
    rng = np.random.RandomState(0)
    for scale_C, shift_d, scale_tau in ((1.0, 0.0, 1.01), (2.0, 1.0, 3.0)):
        pB_np = {'C': scale_C * params['C'] + 0.01 * rng.randn(N, q), 'd': params['d'] + shift_d, 'tau': params['tau'] * scale_tau}
        pB = core.DeviceParams(pB_np['C'], pB_np['d'], pB_np['tau'], T, 10)
        estB = trials.estep_laplace(pB, x0=cold_pcg.x)
        irB, _, _, _ = po.laplace_struct(ys, pB_np, T, 10, want_cov=False)
        assert rel(estB.x, np.stack(irB['post_mean'])) <= 1e-8
        assert rel(estB.vsm, np.stack(irB['post_vsm'])) <= 1e-8
