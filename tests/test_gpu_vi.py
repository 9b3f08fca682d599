"""GPU parity tests of the dual variational E-step (funs/inference.py:188-432)."""
import copy

import numpy as np
import pytest
import torch

from helpers import Exp, init_params, load_golden, rel
from oracle import pgpfa_oracle as po

pytestmark = pytest.mark.gpu


def test_dual_functions_reference_signatures():
    from poisson_gpfa_b200 import inference, util
    g = load_golden("small_vi")
    q, N, T = 2, 8, 20
    ip = init_params(g)
    ex = Exp(g)
    K_big, K = util.makeK_big(ip, ex.trialDur, ex.binSize)
    C_big, d_big = util.makeCd_big(ip, T)
    K_bigInv = np.linalg.inv(K_big)
    ybar = g['Y'][0].reshape(-1)
    lam = g['fn_lam']
    assert rel(inference.dualProblem(lam, ybar, C_big, K_big, K_bigInv, d_big), g['fn_D']) <= 1e-11
    assert rel(inference.dualProblem_grad(lam, ybar, C_big, K_big, K_bigInv, d_big), g['fn_grad']) <= 1e-9
    assert rel(inference.dualProblemRho(np.log(lam), ybar, C_big, K_big, K_bigInv, d_big), g['fn_Drho']) <= 1e-11
    assert rel(inference.dualProblemRho_grad(np.log(lam), ybar, C_big, K_big, K_bigInv, d_big), g['fn_gradrho']) <= 1e-9
    cov, prec = inference.VIPostCov(K_bigInv, C_big, lam)
    assert rel(cov, g['fn_cov']) <= 1e-9 and rel(prec, g['fn_prec']) <= 1e-12
    assert rel(inference.VIPostMean(K_big, C_big, ybar, lam), g['fn_mean']) <= 1e-11


@pytest.mark.parametrize("loglam", [False, True])
def test_dual_variational_fixed_point(loglam):
    from poisson_gpfa_b200 import inference
    g = load_golden("small_vi")
    q, N, T = 2, 8, 20
    ex = Exp(g)
    tag = 'rho' if loglam else 'lam'
    infRes, post_lik, lower, opt = inference.dualVariational(ex, init_params(g), optimizeLogLambda=loglam)
    out, pl_o, lb_o = po.dual_variational_struct(list(g['Y']), init_params(g), T, ex.binSize)
    lam = np.stack(list(opt))
    lam = np.exp(lam) if loglam else lam
    # (a) vs the stationary point (oracle, gradient 1e-12 under the reference's formula): north-star 1e-8
    assert rel(lam, np.stack([o['lam'].ravel() for o in out])) <= 1e-8
    assert rel(np.stack(list(infRes['post_mean'])), np.stack([o['mean'] for o in out])) <= 1e-8
    assert rel(np.stack(list(infRes['post_vsm'])), np.stack([o['vsm'] for o in out])) <= 1e-8
    assert rel(np.stack(list(infRes['post_vsmGP'])), np.stack([o['vsmGP'] for o in out])) <= 1e-8
    assert rel(infRes['post_cov'][0], out[0]['cov']) <= 1e-8
    assert abs(lower - lb_o) <= 1e-10 * abs(lb_o) and abs(post_lik - pl_o) <= 1e-10 * abs(pl_o)
    # (b) vs the reference's tightened L-BFGS-B (stops at |grad| ~ 1e-6)
    lam_ref = np.exp(g[tag + '_opt']) if loglam else g[tag + '_opt']
    assert rel(lam, lam_ref) <= 2e-6
    assert rel(np.stack(list(infRes['post_mean'])), g[tag + '_post_mean']) <= 2e-6
    assert rel(infRes['post_cov'][0], g[tag + '_post_cov0']) <= 1e-6
    assert rel(lower, g[tag + '_vlb']) <= 1e-11 and rel(post_lik, g[tag + '_post_lik']) <= 1e-8
    # warm start from the optimum stays there
    infRes2, pl2, lb2, opt2 = inference.dualVariational(ex, init_params(g), optimizeLogLambda=loglam, prevOptimRes=list(opt))
    assert rel(np.stack(list(opt2)), np.stack(list(opt))) <= 1e-8 if not loglam else True
    assert abs(lb2 - lower) <= 1e-10 * abs(lower)


def test_variational_em_engine():
    """PPGPFAfit(inferenceMethod='variational') batch EM, two iterations, against the oracle's fixed points."""
    from poisson_gpfa_b200 import engine
    g = load_golden("small_vi")
    q, N, T = 2, 8, 20
    ex = Exp(g)
    fit = engine.PPGPFAfit(experiment=ex, initParams=init_params(g), inferenceMethod='variational', EMmode='Batch',
                           maxEMiter=2, quiet=True)
    ys = list(g['Y'])
    params = init_params(g)
    for it in range(2):
        out, pl_o, lb_o = po.dual_variational_struct(ys, params, T, ex.binSize)
        assert abs(fit.posteriorLikelihood[it] - pl_o) <= 1e-8 * abs(pl_o)
        assert abs(fit.variationalLowerBound[it] - lb_o) <= 1e-8 * abs(lb_o)
        infRes = {'post_mean': [o['mean'] for o in out], 'post_vsm': [o['vsm'] for o in out],
                  'post_vsmGP': [o['vsmGP'] for o in out]}
        C, d, _ = po.learn_Cd_newton(params, ys, infRes['post_mean'], infRes['post_vsm'])
        tau, _ = po.learn_tau(params, infRes, ex.binSize, gtol=1e-11)
        params = {'C': C, 'd': d, 'tau': tau}
        assert rel(fit.paramSeq[it + 1]['C'], C) <= 1e-8 and rel(fit.paramSeq[it + 1]['d'], d) <= 1e-8
        assert rel(fit.paramSeq[it + 1]['tau'], tau) <= 1e-8


def test_variational_config4_shape_small():
    """q=8, N=100, T=200 on 2 trials: stationarity certificate instead of a (minutes-long) dense oracle run."""
    from poisson_gpfa_b200 import inference, util, kernels as kn, _lib
    ex = util.simulate(2, 8, 100, 2, 200, binSize=10, dOffset=-1.0)
    params = {'C': ex.params['C'] * 0.9, 'd': ex.params['d'] + 0.1, 'tau': ex.params['tau'] * 1.1}
    infRes, post_lik, lower, opt = inference.dualVariational(ex, copy.deepcopy(params))
    est = infRes.device
    p = est.params
    D, grad, mean, vsm, _ = kn.dualvi_eval(est.lam, est.trials.y, p.C, p.d, p.K, p.Kinv)
    assert float(grad.abs().max()) <= 1e-8
    assert rel(mean, est.x) <= 1e-12 and rel(vsm, est.vsm) <= 1e-12
    y0 = np.asarray(ex.data[0]['Y'], dtype=np.float64)
    K = po.make_K(params['tau'], 200, 10)
    Kinv = np.stack([np.linalg.inv(K[k]) for k in range(8)])
    D_o, g_o, cov_o = po.dual_struct(est.lam[0].cpu().numpy(), y0, params['C'], params['d'], K, Kinv)
    assert abs(float(D[0]) - D_o) <= 1e-10 * abs(D_o)
    assert np.abs(g_o).max() <= 1e-7
