"""The iteration loops run from device-side counters: these tests pin that the control flow does
not change the arithmetic.

* PCG / inexact Newton (funs/inference.py:119-126): the driver enqueues `depth` loop iterations ahead of the last
  count it has read from the progress ring; depth 0 waits for every count (= the host-driven loop of round 1).  Every
  depth must give bit-identical results.
* C,d Newton (funs/learning.py:124-130): pgpfa_mstep_cd_solve (blind schedule, gated kernels) vs the same two entry
  points driven one iteration at a time with a host read in between (round 1's loop): bit-identical.
* Timescale search (funs/learning.py:283-288): pgpfa_mstep_tau_solve vs the numpy search of round 1 driven with the
  device's own cost/gradient evaluations: same evaluations count, p equal to 1e-12 (libm vs CUDA log/pow differ by ulps).
* One EM iteration through DeviceTrials.em_step makes at most 3 host synchronisations and equals the API path.
"""
import copy
import ctypes

import numpy as np
import pytest
import torch

from oracle import pgpfa_oracle as po
from test_gpu_kernels import dev, problem, rel
from test_tau_search_host import host_search_round1

pytestmark = pytest.mark.gpu


def _setup(seed, q, N, T, R):
    from poisson_gpfa_b200 import core, _lib
    ex, ys, params = problem(seed, q, N, T, R)
    trials = core.DeviceTrials(_lib.dev_f64(np.stack(ys)), 10)
    p = core.DeviceParams(params['C'], params['d'], params['tau'], T, 10)
    return ex, ys, params, trials, p


@pytest.mark.parametrize("q,N,T,R", [(3, 7, 40, 9), (8, 100, 200, 24), (2, 20, 50, 70)])
def test_pcg_loop_depth_is_bit_identical(q, N, T, R):
    from poisson_gpfa_b200 import _lib
    ex, ys, params, trials, p = _setup(11 + q, q, N, T, R)
    outs = []
    try:
        for depth in (0, 1, 4, 9):
            _lib.call("pgpfa_set_loop_depth", _lib.handle(), depth)
            est = trials.estep_laplace(p)
            torch.cuda.synchronize()
            outs.append((est.x.clone(), est.f.clone(), est.vsm.clone(), est.vsmGP.clone(), est.niter.clone(),
                         est.stats["pcg_newton_iters"], est.stats["pcg_iters"]))
    finally:
        _lib.call("pgpfa_set_loop_depth", _lib.handle(), 4)
    ref = outs[0]
    assert ref[5] >= 2 and ref[6] >= 4
    for o in outs[1:]:
        for a, b in zip(ref[:5], o[:5]):
            assert torch.equal(a, b)
        assert o[5] == ref[5] and o[6] == ref[6]          # the same iterations had work


@pytest.mark.parametrize("q,N,T,R,pw", [(3, 7, 40, 4, 0.0), (8, 100, 200, 3, 0.0), (2, 20, 50, 5, 6.25)])
def test_cd_solve_equals_host_driven_newton(q, N, T, R, pw):
    from poisson_gpfa_b200 import _lib, kernels as kn
    from poisson_gpfa_b200._lib import call, ptr, stream, empty
    ex, ys, params, trials, p = _setup(5 + q, q, N, T, R)
    est = trials.estep_laplace(p)
    C1, d1, cost1, it1, _ = trials.mstep_cd(p, est, prior_w=pw)          # device-driven (pgpfa_mstep_cd_solve)
    # round 1's loop: one iteration per call, n_open read by the host after each
    P = q + 1
    theta0 = p.theta
    th_cur, th_try = theta0.clone(), theta0.clone()
    fcur, alpha, slope, step = empty(N), empty(N), empty(N), empty(N, P)
    done = torch.zeros(N, dtype=torch.int32, device="cuda")
    n_open = torch.zeros(4, dtype=torch.int32, device="cuda")
    inv_R = 1.0 / trials.R_total
    it = 0
    for it in range(1, 101):
        stats = kn.mstep_cd_stats(trials.y, est.x, est.vsm, th_try)
        call("pgpfa_mstep_cd_update", ptr(stats), inv_R, float(pw), None, ptr(theta0), ptr(th_cur), ptr(th_try), ptr(fcur),
             ptr(step), ptr(alpha), ptr(slope), ptr(done), 1 if it == 1 else 0, 1e-10, N, q, ptr(n_open), it, stream())
        if int(n_open[0].item()) == 0:
            break
    assert it == it1
    assert torch.equal(th_cur[:, :q].contiguous(), C1) and torch.equal(th_cur[:, q].contiguous(), d1)
    assert float(fcur.sum()) == cost1
    C_o, d_o, _ = po.learn_Cd_newton(params, ys, [m.cpu().numpy() for m in est.x], [v.cpu().numpy() for v in est.vsm],
                                     prior_weight=pw)
    assert rel(C1, C_o) <= 1e-8 and rel(d1, d_o) <= 1e-8


@pytest.mark.parametrize("q,N,T,R,prior_step", [(3, 7, 40, 4, None), (8, 100, 200, 3, None), (2, 20, 50, 5, 0.5)])
def test_tau_solve_equals_host_driven_search(q, N, T, R, prior_step):
    from poisson_gpfa_b200 import _lib, kernels as kn, core
    ex, ys, params, trials, p = _setup(7 + q, q, N, T, R)
    est = trials.estep_laplace(p)
    Psum = trials.pautosum(est)
    tau_dev, det = trials.mstep_tau(p, Psum, prior_step=prior_step)
    m = 9
    pw = 0.0 if prior_step is None else 1.0 / prior_step ** 2
    P_rep = Psum.repeat(m, 1, 1).contiguous()
    tau_old_rep = p.tau.repeat(m).contiguous()

    def fg(cands):
        c, g = kn.tau_eval(_lib.dev_f64(np.ascontiguousarray(cands).reshape(-1)), P_rep, float(trials.R_total), T,
                           core.EPS_NOISE, pw, tau_old_rep, 10.0)
        return c.cpu().numpy().reshape(m, q), g.cpu().numpy().reshape(m, q)
    p0 = np.log(1.0 / (np.ravel(params['tau']) * 1000.0 / 10) ** 2)
    p_host, nev_host, _ = host_search_round1(fg, p0)
    assert det['nfev'] == nev_host
    assert np.abs(det['p'] - p_host).max() <= 1e-12
    tau_host = np.sqrt(1.0 / np.exp(p_host)) * 10 / 1000.0
    assert rel(tau_dev, tau_host) <= 1e-12


@pytest.mark.parametrize("q,N,T,R", [(3, 12, 64, 6), (8, 100, 200, 8)])
def test_em_step_host_syncs_and_api_equivalence(q, N, T, R):
    from poisson_gpfa_b200 import _lib, core, inference, learning
    ex, ys, params, trials, p = _setup(21 + q, q, N, T, R)
    newp, est, lik, info = trials.em_step(p)                  # first call: allocations, prior of the first iteration
    s0 = _lib.host_sync_count()
    newp2, est2, lik2, info2 = trials.em_step(newp, x0=est.x)
    syncs = _lib.host_sync_count() - s0
    # one wait inside the E-step driver (fallback count) + one packed read of the M-step flags; the CG / Newton loops
    # only throttle (the device keeps `depth` iterations queued), the M-step loops run blind
    assert syncs <= 3, syncs
    assert 1 <= info2["cd_iters"] <= 6 and 2 <= info2["tau_evals"] <= 6
    # the reference-facing API (separate calls, host dictionaries in between) computes the same iteration
    exp = po.Experiment([{'Y': y} for y in ys], T * 10, 10)
    par = {k: v.copy() for k, v in params.items()}
    infRes, lik_api, optim = inference.laplace(exp, par)
    par2, det = learning.updateParams(par, infRes, exp)
    assert lik_api == lik
    assert np.array_equal(par2['C'], newp.C.cpu().numpy()) and np.array_equal(par2['d'], newp.d.cpu().numpy())
    assert np.array_equal(par2['tau'], newp.tau.cpu().numpy())
    # and it is the oracle's iteration
    p_o, lik_o, _, ir = po.em_step_struct(ys, copy.deepcopy(params), T, 10)
    assert abs(lik - lik_o) <= 1e-10 * abs(lik_o)
    assert rel(newp.C, p_o['C']) <= 1e-8 and rel(newp.d, p_o['d']) <= 1e-8 and rel(newp.tau, p_o['tau']) <= 1e-8


def test_empty_shard_contributes_zero_statistics():
    """A rank whose shard of a mini-batch is empty (batchSize < world) launches nothing and adds zeros."""
    from poisson_gpfa_b200 import core, kernels as kn
    q, N, T = 2, 6, 30
    trials = core.DeviceTrials(torch.zeros(0, N, T, dtype=torch.float64, device="cuda"), 10, R_total=5)
    p = core.DeviceParams(np.zeros((N, q)), np.zeros(N), np.array([0.1, 0.2]), T, 10)
    est = trials.estep_laplace(p)
    assert est.x.shape == (0, q, T) and est.vsm.shape == (0, T, q, q)
    assert float(trials.pautosum(est).abs().max()) == 0.0
    _, g, stats = trials.cd_cost_grad(p.theta, est)
    assert float(stats.abs().max()) == 0.0
    assert float(est.flags().abs().max()) == 0.0
