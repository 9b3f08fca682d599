"""Generates tests/golden/*.npz by running the UNMODIFIED reference (build container only).

Protocol ("parity protocol"): the reference's scipy optimisers are run with tightened
tolerances (oracle/ref_harness.tight_tolerances) so that both sides sit at the fixed point; the
default-tolerance results are stored beside them to show how far the stock reference stops short.
Run: python oracle/make_golden.py
"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_harness as rh     # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def stackY(ds):
    return np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ds.data])


def em_case(ref, name, ds, ip, n_iter, q, N, T):
    """Batch Laplace EM: per-iteration E-step and M-step results with tight tolerances, each M-step
    also evaluated from the reference's own previous parameters (teacher forcing)."""
    out = {'Y': stackY(ds), 'binSize': ds.binSize, 'trialDur': ds.trialDur,
           'init_C': ip['C'], 'init_d': ip['d'], 'init_tau': ip['tau']}
    Kb, K = ref.util.makeK_big(copy.deepcopy(ip), ds.trialDur, ds.binSize)
    out['K0'] = K
    params = copy.deepcopy(ip)
    prev = None
    with rh.tight_tolerances(), rh.quiet():
        for it in range(n_iter):
            infRes, nll, prev = ref.inference.laplace(ds, params, prevOptimRes=prev)
            out['it%d_post_lik' % it] = nll
            out['it%d_post_mean' % it] = np.stack(infRes['post_mean'])
            out['it%d_post_vsm' % it] = np.stack(infRes['post_vsm'])
            out['it%d_post_vsmGP' % it] = np.stack(infRes['post_vsmGP'])
            if it == 0:
                out['it0_post_cov0'] = infRes['post_cov'][0]
                # function-level goldens at the first trial, perturbed point
                rng = np.random.RandomState(7)
                Cb, db = ref.util.makeCd_big(params, T)
                Kinv = np.linalg.inv(Kb)
                x = 0.2 * rng.randn(q * T)
                yb = out['Y'][0].reshape(-1)
                out['fn_x'] = x
                out['fn_f'] = ref.inference.negLogPosteriorUnNorm(x, yb, Cb, db, Kinv, q, N)
                out['fn_g'] = ref.inference.negLogPosteriorUnNorm_grad(x, yb, Cb, db, Kinv, q, N)
                out['fn_H'] = ref.inference.negLogPosteriorUnNorm_hess(x, yb, Cb, db, Kinv, q, N)
                vec = ref.util.CdtoVecCd(params['C'], params['d']) + 0.01 * rng.randn(q * N + N)
                out['fn_vecCd'] = vec
                out['fn_cd_cost'] = ref.learning.MStepObservationCost(vec, q, N, ds, infRes)
                out['fn_cd_grad'] = ref.learning.MStepObservationCost_grad(vec, q, N, ds, infRes)
                Lam = -np.eye(q * N + N) / 0.4 ** 2
                out['fn_cd_cost_prior'] = ref.learning.MStepObservationCostWithPrior(vec, params, q, N, ds, infRes, Lam)
                out['fn_cd_grad_prior'] = ref.learning.MStepObservationCostWithPrior_grad(vec, params, q, N, ds, infRes, Lam)
                pre = ref.learning.makePrecomp(infRes)
                out['fn_PautoSum'] = np.stack([p['PautoSum'] for p in pre])
                pp = np.log(1 / (params['tau'] * 1000 / ds.binSize) ** 2) + 0.15
                out['fn_tau_p'] = pp
                out['fn_tau_cost'] = np.array([ref.learning.MStepGPtimescaleCost(pp[k], pre[k], 0.001) for k in range(q)])
                out['fn_tau_grad'] = np.array([ref.learning.MStepGPtimescaleCost_grad(pp[k], pre[k], 0.001) for k in range(q)])
                out['fn_tau_cost_prior'] = np.array([ref.learning.MStepGPtimescaleCostWithPrior(
                    pp[k], pre[k], 0.001, ds.binSize, params['tau'][k], 0.5) for k in range(q)])
                out['fn_tau_grad_prior'] = np.array([ref.learning.MStepGPtimescaleCostWithPrior_grad(
                    pp[k], pre[k], 0.001, ds.binSize, params['tau'][k], 0.5) for k in range(q)])
            params, det = ref.learning.updateParams(params, infRes, ds, CdOptimMethod='TNC')
            out['it%d_new_C' % it] = params['C']
            out['it%d_new_d' % it] = params['d']
            out['it%d_new_tau' % it] = params['tau']
            out['it%d_cd_cost' % it] = det['Cd']
    # stock tolerances, free-running (what a user of the reference gets), with leave-one-out prediction on the
    # smaller cases (R*N scipy fmin_ncg solves)
    with rh.quiet():
        fit = ref.engine.PPGPFAfit(experiment=ds, initParams=copy.deepcopy(ip), inferenceMethod='laplace',
                                   EMmode='Batch', maxEMiter=n_iter, getPredictionErr=True)
    out['stock_y_pred_mode'] = fit.y_pred_mode
    out['stock_pred_err_mode'] = fit.pred_err_mode
    out['stock_C'] = fit.optimParams['C']
    out['stock_d'] = fit.optimParams['d']
    out['stock_tau'] = fit.optimParams['tau']
    out['stock_post_lik'] = np.array(fit.posteriorLikelihood)
    out['n_iter'] = n_iter
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'written;', {k: np.shape(v) for k, v in list(out.items())[:3]})


def online_case(ref, name, ds, ip, n_iter, batchSize, method):
    """Online EM ('diag' rule) with tight tolerances: batch indices, per-iteration params."""
    out = {'Y': stackY(ds), 'binSize': ds.binSize, 'trialDur': ds.trialDur,
           'init_C': ip['C'], 'init_d': ip['d'], 'init_tau': ip['tau'], 'batchSize': batchSize}
    np.random.seed(2024)
    with rh.tight_tolerances(), rh.quiet():
        fit = ref.engine.PPGPFAfit(experiment=ds, initParams=copy.deepcopy(ip), inferenceMethod='laplace',
                                   EMmode='Online', maxEMiter=n_iter, batchSize=batchSize,
                                   onlineParamUpdateMethod=method)
    # replay the RNG to record the batches the reference drew
    np.random.seed(2024)
    out['batches'] = np.stack([np.random.choice(len(ds.data), batchSize, replace=False) for _ in range(n_iter)])
    out['seq_C'] = np.stack([p['C'] for p in fit.paramSeq])
    out['seq_d'] = np.stack([p['d'] for p in fit.paramSeq])
    out['seq_tau'] = np.stack([np.ravel(p['tau']) for p in fit.paramSeq])
    out['post_lik'] = np.array(fit.posteriorLikelihood)
    out['seed'] = 2024
    if method in ('hess', 'diag'):
        out['invPriorCov_last'] = fit.invPriorCovs[-1]
        # iteration 1 of 'hess': minus the reference's finite-difference Jacobian (funs/util.py:377-434) of the prior-cost
        # gradient at the initial parameters (funs/learning.py:545-549) — the function-level pin of the analytic blocks
        out['invPriorCov_1'] = fit.invPriorCovs[1]
    if method == 'grad':
        out['cumHess_last'] = fit.cumHess[-1]
        out['cumHess_1'] = fit.cumHess[1]          # I + the FD Hessian of the first mini-batch (funs/learning.py:884-891)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'written')


def vi_case(ref, name, ds, ip, q, N, T):
    """Dual variational E-step: function values at a random lambda and the reference's tightened optimum
    (both the bounded-lambda and the log-lambda variants), first E-step only."""
    out = {'Y': stackY(ds), 'binSize': ds.binSize, 'trialDur': ds.trialDur,
           'init_C': ip['C'], 'init_d': ip['d'], 'init_tau': ip['tau']}
    params = copy.deepcopy(ip)
    Kb, K = ref.util.makeK_big(params, ds.trialDur, ds.binSize)
    Cb, db = ref.util.makeCd_big(params, T)
    Kinv = np.linalg.inv(Kb)
    rng = np.random.RandomState(11)
    lam = np.exp(0.4 * rng.randn(N * T))
    yb = out['Y'][0].reshape(-1)
    out['fn_lam'] = lam
    out['fn_D'] = ref.inference.dualProblem(lam, yb, Cb, Kb, Kinv, db)
    out['fn_grad'] = ref.inference.dualProblem_grad(lam, yb, Cb, Kb, Kinv, db)
    out['fn_Drho'] = ref.inference.dualProblemRho(np.log(lam), yb, Cb, Kb, Kinv, db)
    out['fn_gradrho'] = ref.inference.dualProblemRho_grad(np.log(lam), yb, Cb, Kb, Kinv, db)
    cov, prec = ref.inference.VIPostCov(Kinv, Cb, lam)
    out['fn_cov'], out['fn_prec'] = cov, prec
    out['fn_mean'] = ref.inference.VIPostMean(Kb, Cb, yb, lam)
    for tag, loglam in (('lam', False), ('rho', True)):
        with rh.tight_tolerances(), rh.quiet():
            infRes, nll, vlb, opt = ref.inference.dualVariational(ds, copy.deepcopy(ip), optimizeLogLambda=loglam)
        out[tag + '_opt'] = np.stack(opt)
        out[tag + '_post_mean'] = np.stack(infRes['post_mean'])
        out[tag + '_post_vsm'] = np.stack(infRes['post_vsm'])
        out[tag + '_post_vsmGP'] = np.stack(infRes['post_vsmGP'])
        out[tag + '_post_cov0'] = infRes['post_cov'][0]
        out[tag + '_post_lik'] = nll
        out[tag + '_vlb'] = vlb
        # gradient of the dual at the reference's own optimum: how far scipy's L-BFGS-B stops from stationarity
        lam_star = np.exp(opt[0]) if loglam else opt[0]
        out[tag + '_grad_at_opt0'] = ref.inference.dualProblem_grad(lam_star, yb, Cb, Kb, Kinv, db)
    np.savez_compressed(os.path.join(OUT, name + '.npz'), **out)
    print(name, 'written')


def main():
    os.makedirs(OUT, exist_ok=True)
    ref = rh.load_reference()
    # config 1: example.py verbatim (seed 123, q=2, N=20, R=5, T=50)
    np.random.seed(123)
    with rh.quiet():
        ds = ref.util.dataset(seed=np.random.randint(10000), xdim=2, ydim=20, numTrials=5, trialDur=1000, binSize=20,
                              dOffset=1, fixTau=True, fixedTau=np.linspace(0.1, 0.5, 2), drawSameX=True)
        ip = ref.util.initializeParams(2, 20, ds)
    only = [a.split('=', 1)[1].split(',') for a in sys.argv if a.startswith('--only=')]
    only = only[0] if only else None              # e.g. --only=example_online_hess,example_online_grad
    want = lambda name: only is None or name in only
    if want('example_laplace'):
        em_case(ref, 'example_laplace', ds, ip, 3, 2, 20, 50)
    for nm, it_, meth in (('example_online_diag', 4, 'diag'), ('example_online_hess', 3, 'hess'), ('example_online_grad', 3, 'grad')):
        if want(nm):
            online_case(ref, nm, ds, ip, it_, 3, meth)
    if only is not None:
        return
    # small ragged-ish shape: q=3, N=7 (few neurons), T=40, tile-unaligned n=120
    np.random.seed(5)
    with rh.quiet():
        ds2 = ref.util.dataset(seed=77, xdim=3, ydim=7, numTrials=6, trialDur=400, binSize=10, dOffset=0.5,
                               fixTau=True, fixedTau=np.linspace(0.04, 0.15, 3))
        ip2 = ref.util.initializeParams(3, 7, ds2)
    em_case(ref, 'small_q3_laplace', ds2, ip2, 2, 3, 7, 40)
    # variational: q=2, N=8, T=20, 3 trials (each reference gradient forms an NT x NT matrix)
    if 'vi' in sys.argv or len(sys.argv) == 1:
        np.random.seed(9)
        with rh.quiet():
            ds3 = ref.util.dataset(seed=31, xdim=2, ydim=8, numTrials=3, trialDur=200, binSize=10, dOffset=0.5,
                                   fixTau=True, fixedTau=np.linspace(0.05, 0.12, 2))
            ip3 = ref.util.initializeParams(2, 8, ds3)
        vi_case(ref, 'small_vi', ds3, ip3, 2, 8, 20)


if __name__ == "__main__":
    main()
