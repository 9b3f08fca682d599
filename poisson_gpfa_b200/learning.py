"""Mirror of the hot-path part of the reference's ``funs/learning.py`` (M-step) on the B200 kernels.

C,d: the reference hands MStepObservationCost to scipy (TNC / BFGS / L-BFGS-B ...,
funs/learning.py:124-130, :605-622).  The cost is a sum over neurons of convex (q+1)-dimensional
problems, so the rebuild runs an exact per-neuron Newton on the device; ``CdOptimMethod`` is accepted
for signature compatibility.  The result is the fixed point all of those methods approach (scipy's
line-search methods stall ~1e-7 short of it, DESIGN.md "parity protocol").
tau: q scalar problems, bracket + safeguarded secant on the reference's own gradient.
"""
import numpy as np
import torch

from . import _lib, kernels as kn, util
from .core import EPS_NOISE, DeviceParams
from .inference import InfRes, as_estep_result, device_params, device_trials

_f64 = _lib.dev_f64


class OptimizeDetail(dict):
    """Small stand-in for scipy's OptimizeResult (attribute access to the same keys)."""
    __getattr__ = dict.get


def _theta_from_vec(vecCd, xdim, ydim):
    C, d = util.vecCdtoCd(np.asarray(vecCd, dtype=np.float64), xdim, ydim)
    return _f64(np.concatenate([C, d[:, None]], axis=1))


# ---------------------------------------------------------------------------------------------- C, d
def MStepObservationCost(vecCd, xdim, ydim, experiment, infRes):
    """funs/learning.py:20-49."""
    est = as_estep_result(infRes, experiment)
    cost, _, _ = est.trials.cd_cost_grad(_theta_from_vec(vecCd, xdim, ydim), est)
    return cost


def MStepObservationCost_grad(vecCd, xdim, ydim, experiment, infRes):
    """funs/learning.py:51-91 (returned in CdtoVecCd layout)."""
    est = as_estep_result(infRes, experiment)
    _, g, _ = est.trials.cd_cost_grad(_theta_from_vec(vecCd, xdim, ydim), est)
    g = g.cpu().numpy()
    return util.CdtoVecCd(g[:, :xdim], g[:, xdim])


def MStepObservationCostWithPrior(vecCd, oldParams, xdim, ydim, experiment, infRes, invPriorCov):
    """funs/learning.py:445-486: base cost - 0.5 (v-v_old)^T invPriorCov (v-v_old)."""
    dv = np.asarray(vecCd) - util.CdtoVecCd(oldParams['C'], oldParams['d'])
    return MStepObservationCost(vecCd, xdim, ydim, experiment, infRes) - 0.5 * float(dv @ (np.asarray(invPriorCov) @ dv))


def MStepObservationCostWithPrior_grad(vecCd, oldParams, xdim, ydim, experiment, infRes, invPriorCov):
    """funs/learning.py:488-534."""
    dv = np.asarray(vecCd) - util.CdtoVecCd(oldParams['C'], oldParams['d'])
    return MStepObservationCost_grad(vecCd, xdim, ydim, experiment, infRes) - np.asarray(invPriorCov) @ dv


def learnLTparams(oldParams, infRes, experiment, CdOptimMethod='TNC', CdMaxIter=None, verbose=False, tol=1e-10):
    """newC, newd, cost = learnLTparams(...) — funs/learning.py:93-141."""
    est = as_estep_result(infRes, experiment)
    p = device_params(oldParams, est.trials.T, experiment.binSize)
    C, d, cost, it, _ = est.trials.mstep_cd(p, est, tol=tol, max_iter=CdMaxIter or 100)
    if verbose:
        print('Cd optimization: %d Newton iterations, cost %.10g' % (it, cost))
    return C.cpu().numpy(), d.cpu().numpy(), cost


def _dense_hessian_from_stats(stats, q, N, scale):
    """(Nq+N)^2 Hessian in CdtoVecCd layout from per-neuron packed upper blocks."""
    P = q + 1
    st = stats.cpu().numpy()
    iu = np.triu_indices(P)
    H = np.zeros((N * q + N, N * q + N))
    pos = lambda k, n: (k * N + n)          # vec index of theta[n,k] (k == q -> d)
    for b, (k, l) in enumerate(zip(*iu)):
        v = st[1 + P + b] * scale
        for n in range(N):
            H[pos(k, n), pos(l, n)] = v[n]
            H[pos(l, n), pos(k, n)] = v[n]
    return H


def learnLTparamsWithPrior(oldParams, infRes, experiment, CdOptimMethod, regularizer_stepsize_Cd, prevInvPriorCov,
                           covOpts='useDiag', updateCdJointly=True, hessTol=1e-5, verbose=False, tol=1e-10):
    """newC, newd, cost, invPriorCov — funs/learning.py:536-676.
    'useDiag': proximal term I/s^2 (the default online rule, funs/engine.py:370-385).
    'useHessian': the reference builds invPriorCov from a finite-difference Jacobian of the
    prior-cost gradient (4(Nq+N) gradient passes, :545-549); here the analytic block Hessian of the
    same function is used (FD accuracy ~1e-7 relative) and the per-neuron blocks
    keep the problem separable."""
    if not updateCdJointly:
        raise NotImplementedError("updateCdJointly=False fails in the reference itself on current numpy "
                                  "(funs/learning.py:393, ValueError) and is not supported")
    est = as_estep_result(infRes, experiment)
    trials = est.trials
    N, q = np.shape(oldParams['C'])
    p = device_params(oldParams, trials.T, experiment.binSize)
    if covOpts == 'useDiag':
        pw = 1.0 / float(regularizer_stepsize_Cd) ** 2
        invPriorCov = -np.diag(np.ones(q * N + N)) / (regularizer_stepsize_Cd ** 2)
        C, d, cost, it, _ = trials.mstep_cd(p, est, prior_w=pw, tol=tol)
    elif covOpts == 'useHessian':
        # reference: invPriorCov = -J, J = finite-difference Jacobian at the old parameters of the prior-cost gradient
        # = H_base(theta_old) - prevInvPriorCov  (funs/learning.py:545-549).  H_base is block diagonal over neurons;
        # the per-neuron blocks of prevInvPriorCov are used (the reference's own matrices are block diagonal up to FD noise).
        P = q + 1
        pos = (np.arange(P)[:, None] * N + np.arange(N)[None, :])            # vec index of theta[n,k]: pos[k,n]
        prev = np.asarray(prevInvPriorCov, dtype=np.float64)
        prev_blocks = np.stack([prev[np.ix_(pos[:, n], pos[:, n])] for n in range(N)])          # (N,P,P)
        _, _, stats = trials.cd_cost_grad(p.theta, est)
        st = stats.cpu().numpy() / trials.R_total
        iu = np.triu_indices(P)
        Hb = np.zeros((N, P, P))
        Hb[:, iu[0], iu[1]] = st[1 + P:].T
        Hb[:, iu[1], iu[0]] = st[1 + P:].T
        lam_blocks = prev_blocks - Hb
        pmat = _f64(np.ascontiguousarray((-lam_blocks)[:, iu[0], iu[1]].T))                      # (P(P+1)/2, N)
        C, d, cost, it, _ = trials.mstep_cd(p, est, prior_mat=pmat, tol=tol)
        invPriorCov = np.zeros((q * N + N, q * N + N))
        for n in range(N):
            invPriorCov[np.ix_(pos[:, n], pos[:, n])] = lam_blocks[n]
    else:
        raise ValueError("covOpts must be 'useDiag' or 'useHessian'")
    if verbose:
        print('Cd optimization with prior: %d Newton iterations' % it)
    return C.cpu().numpy(), d.cpu().numpy(), cost, invPriorCov


def learnLTparamsGradDescent(oldParams, infRes, experiment, stepSize, cumHess, updateCdJointly=True, hessTol=1e-5):
    """newC, newd, h — funs/learning.py:875-907: one Newton step vecCd -= stepSize * inv(h) g of
    Q = -MStepObservationCost.  The reference differentiates the gradient numerically
    (util.approx_jacobian); here h is the analytic Hessian of the same Q."""
    if not updateCdJointly:
        raise NotImplementedError("updateCdJointly=False is not supported (see learnLTparamsWithPrior)")
    est = as_estep_result(infRes, experiment)
    N, q = np.shape(oldParams['C'])
    p = device_params(oldParams, est.trials.T, experiment.binSize)
    C, d, _, _, stats = est.trials.mstep_cd(p, est, one_step=True, step_size=float(stepSize))
    h = -_dense_hessian_from_stats(stats, q, N, 1.0 / est.trials.R_total)
    return C.cpu().numpy(), d.cpu().numpy(), h


# ---------------------------------------------------------------------------------------------- tau
def makePrecomp(infRes):
    """funs/learning.py:145-173 (list over latents of dicts with T, Tdif, difSq, numTrials, PautoSum)."""
    if isinstance(infRes, InfRes):
        est = infRes.device
        P = est.trials.pautosum(est).cpu().numpy()
        R = est.trials.R_total
    else:
        x = _f64(np.stack([np.asarray(m) for m in infRes['post_mean']]))
        vs = _f64(np.stack([np.asarray(v).transpose(2, 0, 1) for v in infRes['post_vsmGP']]))
        P = kn.pautosum(vs, x).cpu().numpy()
        R = x.shape[0]
    q, T = P.shape[0], P.shape[1]
    idx = np.arange(T) + 1
    Tdif = idx[:, None] - idx[None, :]
    return [{'T': T, 'Tdif': Tdif, 'difSq': Tdif * Tdif, 'numTrials': R, 'PautoSum': P[k]} for k in range(q)]


def _tau_eval_single(p, precomp, epsNoise, prior=None):
    T = precomp['T']
    P = _f64(precomp['PautoSum'][None])
    pv = _f64(np.atleast_1d(np.asarray(p, dtype=np.float64))[:1])
    if prior is None:
        return kn.tau_eval(pv, P, precomp['numTrials'], T, epsNoise)
    binSize, oldTau, step = prior
    return kn.tau_eval(pv, P, precomp['numTrials'], T, epsNoise, 1.0 / float(step) ** 2,
                       _f64(np.atleast_1d(oldTau)), binSize)


def MStepGPtimescaleCost(p, precomp, epsNoise):
    """funs/learning.py:175-214."""
    return float(_tau_eval_single(p, precomp, epsNoise)[0][0])


def MStepGPtimescaleCost_grad(p, precomp, epsNoise):
    """funs/learning.py:216-255."""
    return float(_tau_eval_single(p, precomp, epsNoise)[1][0])


def MStepGPtimescaleCostWithPrior(p, precomp, epsNoise, binSize, oldTau, regularizer_stepsize_tau):
    """funs/learning.py:681-724."""
    return float(_tau_eval_single(p, precomp, epsNoise, (binSize, oldTau, regularizer_stepsize_tau))[0][0])


def MStepGPtimescaleCostWithPrior_grad(p, precomp, epsNoise, binSize, oldTau, regularizer_stepsize_tau):
    """funs/learning.py:726-769."""
    return float(_tau_eval_single(p, precomp, epsNoise, (binSize, oldTau, regularizer_stepsize_tau))[1][0])


def _learn_tau(oldParams, infRes, experiment, prior_step=None):
    est = as_estep_result(infRes, experiment)
    trials = est.trials
    p = device_params(oldParams, trials.T, experiment.binSize)
    Psum = trials.pautosum(est)
    tau, det = trials.mstep_tau(p, Psum, prior_step=prior_step)
    tau = _lib.to_host(tau)
    details = [OptimizeDetail(x=np.array([det['p'][k]]), fun=det['fun'][k], jac=np.array([det['grad'][k]]),
                              nfev=det['nfev'], success=bool(det['bracketed'][k])) for k in range(p.q)]
    return tau, details


def learnGPparams(oldParams, infRes, experiment):
    """newTau, details — funs/learning.py:257-293."""
    return _learn_tau(oldParams, infRes, experiment)


def learnGPparamsWithPrior(oldParams, infRes, experiment, tauOptimMethod, regularizer_stepsize_tau):
    """funs/learning.py:771-830."""
    return _learn_tau(oldParams, infRes, experiment, prior_step=regularizer_stepsize_tau)


# ---------------------------------------------------------------------------------------------- glue
def updateParams(oldParams, infRes, experiment, CdOptimMethod='BFGS', CdMaxIter=None, tauMaxIter=None, verbose=False):
    """(newParams, {'Cd': cost, 'tau': [details]}) — funs/learning.py:295-309.

    learnLTparams and learnGPparams are enqueued together (the C,d Newton iterations on the stream where the posterior
    means were final first, the timescale search behind the covariance sums), the prior of the NEXT E-step is built on
    the device from the new parameters, and ONE packed read brings back the new parameters, the costs and every flag."""
    from .core import read_packed
    from .inference import remember_params
    est = as_estep_result(infRes, experiment)
    trials = est.trials
    p = device_params(oldParams, trials.T, experiment.binSize)
    q = p.q
    cd = trials.mstep_cd_async(p, est, n_blind=min(4, CdMaxIter or 4))
    Psum = trials.pautosum(est)
    ts = trials.mstep_tau_async(p, Psum, n_blind=trials._tau_blind)
    cd.join()
    newp = DeviceParams(cd.th_cur[:, :q].contiguous(), cd.th_cur[:, q].contiguous(), ts.tau, trials.T, experiment.binSize)
    pend_cd, pend_ts, pend_p = cd.pending(), ts.pending(), newp.pending()
    cd.join()
    vals = read_packed(pend_cd + pend_ts + pend_p + [newp.C, newp.d, newp.tau, ts.det])
    v_cd, v_ts, v_p = vals[:3], vals[3:4], vals[4:4 + len(pend_p)]
    Ch, dh, tauh, deth = vals[4 + len(pend_p):]
    redo = v_cd[0][0] > 0 or v_ts[0][0] > 0
    C, d, cost, it, _ = cd.finish(v_cd, max_iter=CdMaxIter or 100)
    tau = ts.finish(v_ts)
    trials._tau_blind = max(2, min(4, int(ts.flags_host[1])))
    if not est._checked:
        from .core import EStepResult
        EStepResult.raise_on(cd.extra_host)
        est._checked = True
    if redo:          # a blind schedule was too short (rare): finish on the host's say-so and read the final values
        Ch, dh, tauh, deth = read_packed([C, d, tau, ts.det])
        newp = DeviceParams(C, d, tau, trials.T, experiment.binSize)
    else:
        newp.resolve(v_p)
    new = {'C': Ch, 'd': dh, 'tau': tauh}
    remember_params(new, newp)
    fl = ts.flags_host
    det = {'p': deth[0], 'p0': deth[1], 'grad': deth[2], 'fun': deth[3], 'nfev': int(fl[1]),
           'bracketed': np.array([(int(fl[2]) >> k) & 1 for k in range(q)], dtype=bool)}
    details = [OptimizeDetail(x=np.array([det['p'][k]]), fun=det['fun'][k], jac=np.array([det['grad'][k]]),
                              nfev=det['nfev'], success=bool(det['bracketed'][k])) for k in range(q)]
    if verbose:
        print('Cd optimization: %d Newton iterations, cost %.10g; tau search: %d evaluations' % (it, cost, det['nfev']))
    return new, {'Cd': cost, 'tau': details}


def updateParamsWithPrior(oldParams, infRes, experiment, CdOptimMethod, tauOptimMethod, regularizer_stepsize_Cd,
                          regularizer_stepsize_tau, prevInvPriorCov, covOpts='useHessian', verbose=False,
                          updateCdJointly=True, hessTol=1e-5):
    """(newParams, optimDetails, invPriorCov) — funs/learning.py:833-866."""
    newC, newd, cost, invPriorCov = learnLTparamsWithPrior(
        oldParams, infRes, experiment, CdOptimMethod, regularizer_stepsize_Cd, prevInvPriorCov, covOpts,
        updateCdJointly, hessTol, verbose)
    newTau, det = learnGPparamsWithPrior(oldParams, infRes, experiment, tauOptimMethod, regularizer_stepsize_tau)
    return {'C': newC, 'd': newd, 'tau': newTau}, {'Cd': cost, 'tau': det}, invPriorCov


def updateParamsWithGradDescent(oldParams, infRes, experiment, stepSize, cumHess, regularizer_stepsize_tau,
                                tauOptimMethod, updateCdJointly=True, verbose=False, hessTol=1e-5):
    """(newParams, optimDetails, hess) — funs/learning.py:932-966."""
    newC, newd, hess = learnLTparamsGradDescent(oldParams, infRes, experiment, stepSize, cumHess, updateCdJointly, hessTol)
    newTau, det = learnGPparamsWithPrior(oldParams, infRes, experiment, tauOptimMethod, regularizer_stepsize_tau)
    return {'C': newC, 'd': newd, 'tau': newTau}, {'Cd': None, 'tau': det}, hess
