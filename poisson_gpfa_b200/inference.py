"""Mirror of the reference's ``funs/inference.py`` on the B200 kernels (same names, arguments and
return structures; all arithmetic in libpgpfa_b200.so, no CPU fallback).

``laplace`` / ``dualVariational`` return ``infRes`` dictionaries with the reference's keys
(funs/inference.py:176-180).  The per-trial lists are lazy views over device tensors: they
materialise numpy arrays on access, while ``learning.updateParams`` consumes the device tensors
directly.  ``post_cov`` (qT x qT per trial, 21 GB at the 1024-trial shape) is computed per trial on
demand from the mode.
"""
import numpy as np
import torch

from . import _lib, kernels as kn
from .core import DeviceParams, DeviceTrials, EStepResult
from .dist import Reducer, shard_bounds

_f64 = _lib.dev_f64


# ----------------------------------------------------------------------------------------------
# helpers: experiments <-> device
# ----------------------------------------------------------------------------------------------
def _host_rows(experiment, lo, hi):
    """Rows [lo, hi) of the spike counts as a host array / pinned tensor (no copy of the other trials)."""
    Y_all = getattr(experiment, 'Y_all', None)
    if Y_all is not None:          # optional fast path: all counts as one (R,N,T) array or pinned tensor
        return torch.as_tensor(Y_all)[lo:hi]
    if hi <= lo:                   # a rank without trials (fewer trials than ranks)
        N, T = np.shape(experiment.data[0]['Y']) if len(experiment.data) else (0, 0)
        return torch.zeros(0, N, T, dtype=torch.float64)
    return torch.from_numpy(np.stack([np.asarray(experiment.data[r]['Y'], dtype=np.float64) for r in range(lo, hi)]))


def _full_counts(experiment):
    """All trials of an experiment as one (R,N,T) float64 device tensor, uploaded once and cached on the experiment
    object.  Only needed for parents of mini-batches (util.subsampleTrials), which are gathered on the device."""
    cached = experiment.__dict__.get('_pgpfa_y')
    if cached is None or cached[1] is not experiment.data:
        y = _to_device_f64(_host_rows(experiment, 0, len(experiment.data)))
        cached = (y, experiment.data)
        experiment.__dict__['_pgpfa_y'] = cached
    return cached[0]


def _to_device_f64(rows):
    """Host rows (any numeric dtype; pinned or pageable) -> float64 on the device.  Integer counts travel in their own
    (compact) dtype and are widened on the device: the reference promotes counts to float64 on use as well."""
    if rows.dtype == torch.float64:
        return rows.to(device="cuda", non_blocking=True).contiguous()
    return rows.to(device="cuda", non_blocking=True).to(torch.float64).contiguous()


def upload_counts(experiment):
    """Re-copy this rank's host spike counts into the resident device buffer (keeps workspaces and kept factors)."""
    dt = experiment.__dict__.get('_pgpfa_dev')
    if dt is None:
        return device_trials(experiment).y
    rows = _host_rows(experiment, dt.offset, dt.offset + dt.R)
    if rows.dtype == torch.float64:
        dt.y.copy_(rows, non_blocking=True)
    else:
        dt.y.copy_(rows.to(device="cuda", non_blocking=True))        # H2D in the compact dtype, widened on the device
    return dt.y


def device_trials(experiment, reducer=None):
    """This rank's shard of the experiment (contiguous block of trials).  Batch mode uploads only the
    shard; a mini-batch made by util.subsampleTrials is gathered on the device from its (fully resident) parent and
    then split across the ranks."""
    reducer = reducer if reducer is not None else Reducer()
    R = len(experiment.data)
    cached = experiment.__dict__.get('_pgpfa_dev')
    # the cache is tied to the very list of trials it was built from (a re-assigned experiment.data, even of the same
    # length, is a different data set)
    if (cached is not None and cached.R_total == R and cached.reducer.world_size == reducer.world_size
            and getattr(cached, '_data_ref', None) is experiment.data):
        return cached
    lo, hi = shard_bounds(R, reducer.world_size, reducer.rank)
    parent = experiment.__dict__.get('_pgpfa_parent')
    if parent is not None and hasattr(experiment, 'batchTrIdx'):
        idx = torch.as_tensor(np.asarray(experiment.batchTrIdx)[lo:hi], device="cuda", dtype=torch.long)
        y = _full_counts(parent).index_select(0, idx).contiguous()
    else:
        y = _to_device_f64(_host_rows(experiment, lo, hi))
    if parent is not None and hasattr(experiment, 'batchTrIdx') and y.shape[0] == 0:
        y = y.reshape(0, *_full_counts(parent).shape[1:])
    dt = DeviceTrials(y, experiment.binSize, reducer, R_total=R, offset=lo)
    dt._data_ref = experiment.data
    experiment.__dict__['_pgpfa_dev'] = dt
    return dt


# Device-side parameter sets whose prior (K, K^-1, low-rank factor) is already built, keyed by the parameter VALUES:
# learning.updateParams leaves the set it returns here, so the next inference.laplace(experiment, thoseParams) finds the
# prior of its E-step ready (built on the device underneath the M-step's host read) instead of rebuilding it.
_prepared = {}


def _params_key(C, d, tau, T, binSize):
    h = lambda a: hash(np.ascontiguousarray(np.asarray(a, dtype=np.float64)).tobytes())
    return (int(T), float(binSize), np.shape(C), h(C), h(d), h(tau))


def remember_params(params_host, dparams):
    _prepared.clear()              # one entry: the parameters of the iteration in flight
    _prepared[_params_key(params_host['C'], params_host['d'], params_host['tau'], dparams.T, dparams.binSize)] = dparams


def device_params(params, T, binSize):
    hit = _prepared.get(_params_key(params['C'], params['d'], params['tau'], T, binSize)) if _prepared else None
    if hit is not None:
        return hit
    return DeviceParams(params['C'], params['d'], params['tau'], T, binSize)


class _TrialView:
    """Sequence of per-trial numpy arrays backed by one device tensor (leading axis = trial)."""

    def __init__(self, tensor, fn=None):
        self.tensor, self._fn = tensor, fn

    def __len__(self):
        return self.tensor.shape[0]

    def __getitem__(self, r):
        if isinstance(r, slice):
            return [self[i] for i in range(*r.indices(len(self)))]
        a = self.tensor[r]
        if self._fn is not None:
            a = self._fn(a)
        return a.cpu().numpy()

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class _LazyTrialView(_TrialView):
    """Per-trial view over a device tensor that is only produced on first access (post_vsmGP)."""

    def __init__(self, make, length, fn=None):
        self._make, self._len, self._fn, self._t = make, length, fn, None

    @property
    def tensor(self):
        if self._t is None:
            self._t = self._make()
        return self._t

    def __len__(self):
        return self._len


class _CovView:
    """post_cov[r]: full qT x qT posterior covariance of trial r, built on demand on the device."""

    def __init__(self, est, diag_scale=1.0, W_fn=None):
        self._est, self._ds, self._W_fn = est, diag_scale, W_fn

    def __len__(self):
        return self._est.x.shape[0]

    def __getitem__(self, r):
        if isinstance(r, slice):
            return [self[i] for i in range(*r.indices(len(self)))]
        est = self._est
        p = est.params
        q, T = p.q, p.T
        if self._W_fn is None:
            _, _, W = kn.laplace_eval(est.x[r:r + 1], est.trials.y[r:r + 1], p.C, p.d, p.Kinv)
        else:
            W = self._W_fn(r)
        L, D, ZT, info = kn.potrf_posterior(p.Kinv, W, self._ds)
        kn.trtri(L, D, ZT, q * T)
        return kn.potri_dense(ZT, q * T)[0].cpu().numpy()

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class InfRes(dict):
    """The reference's infRes dict plus the device-resident E-step result (``.device``)."""

    def __init__(self, est, diag_scale=1.0, W_fn=None):
        super().__init__()
        self.device = est
        self['post_mean'] = _TrialView(est.x)
        self['post_cov'] = _CovView(est, diag_scale, W_fn)
        self['post_vsm'] = _TrialView(est.vsm)
        self['post_vsmGP'] = _LazyTrialView(est.get_vsmGP, est.x.shape[0], lambda a: a.permute(1, 2, 0).contiguous())


def as_estep_result(infRes, experiment, params=None):
    """Device view of an infRes: ours directly, or upload a plain reference-style dict of numpy lists."""
    if isinstance(infRes, InfRes):
        return infRes.device
    trials = device_trials(experiment)
    x = _f64(np.stack([np.asarray(m) for m in infRes['post_mean']]))
    vsm = _f64(np.stack([np.asarray(v) for v in infRes['post_vsm']]))
    vsmGP = None
    if 'post_vsmGP' in infRes and infRes['post_vsmGP'] is not None:
        vsmGP = _f64(np.stack([np.asarray(v).transpose(2, 0, 1) for v in infRes['post_vsmGP']]))
    return EStepResult(x, None, vsm, vsmGP, None, None, None, trials)


def _unpack_big(C_big, d_big, K_bigInv, xdim, ydim):
    """Recover C (N,q), d (N), Kinv (q,T,T) from the reference's big matrices (funs/util.py:594-597)."""
    C_big, d_big = np.asarray(C_big), np.asarray(d_big)
    T = int(len(d_big) / ydim)
    C = C_big[::T, ::T].T.copy()                  # C_big[k*T, n*T] = C[n,k]
    d = d_big[::T].copy()
    Kb = np.asarray(K_bigInv)
    Kinv = np.stack([Kb[k * T:(k + 1) * T, k * T:(k + 1) * T] for k in range(xdim)])
    return C, d, Kinv, T


def _eval_big(xbar, ybar, C_big, d_big, K_bigInv, xdim, ydim):
    C, d, Kinv, T = _unpack_big(C_big, d_big, K_bigInv, xdim, ydim)
    x = _f64(np.asarray(xbar, dtype=np.float64).reshape(1, xdim, T))
    y = _f64(np.asarray(ybar, dtype=np.float64).reshape(1, ydim, T))
    Kd = _f64(Kinv)
    f, g, W = kn.laplace_eval(x, y, _f64(C), _f64(d), Kd)
    return f, g, W, Kd


# ----------------------------------------------------------------------------------------------
# Laplace inference
# ----------------------------------------------------------------------------------------------
def negLogPosteriorUnNorm(xbar, ybar, C_big, d_big, K_bigInv, xdim, ydim):
    """funs/inference.py:12-32."""
    f, _, _, _ = _eval_big(xbar, ybar, C_big, d_big, K_bigInv, xdim, ydim)
    return float(f[0])


def negLogPosteriorUnNorm_grad(xbar, ybar, C_big, d_big, K_bigInv, xdim, ydim):
    """funs/inference.py:34-48."""
    _, g, _, _ = _eval_big(xbar, ybar, C_big, d_big, K_bigInv, xdim, ydim)
    return g.reshape(-1).cpu().numpy()


def negLogPosteriorUnNorm_hess(xbar, ybar, C_big, d_big, K_bigInv, xdim, ydim):
    """funs/inference.py:50-65."""
    _, _, W, Kd = _eval_big(xbar, ybar, C_big, d_big, K_bigInv, xdim, ydim)
    return kn.hessian_dense(Kd, W)[0].cpu().numpy()


def laplace(experiment, params, prevOptimRes=None, returnOptimRes=True, verbose=False, optimMethod='Newton-CG',
            tol=1e-8, reducer=None):
    """laplaceInfRes, -post_lik[, lapOptimRes] = laplace(experiment, params) — funs/inference.py:67-185.

    The per-trial scipy Newton-CG loop is replaced by one batched exact-Newton solve on the device
    (``optimMethod`` is accepted for signature compatibility).  ``tol`` bounds the last Newton step
    (relative, inf-norm); the returned mode is quadratically closer than that, and the covariances (evaluated one
    step earlier) are within ~tol of the converged ones.  Failures (a Hessian that is not positive definite, the
    iteration limit, a non-finite objective) raise here, on every rank of a sharded run."""
    trials = device_trials(experiment, reducer)
    T = trials.T
    p = device_params(params, T, experiment.binSize)
    params['tau'] = np.ndarray.flatten(np.asarray(params['tau'], dtype=np.float64))   # funs/util.py:602 side effect
    x0 = None
    if prevOptimRes is not None:
        if isinstance(prevOptimRes, _TrialView) and prevOptimRes.tensor.shape[0] == trials.R:
            x0 = prevOptimRes.tensor.reshape(trials.R, p.q, T)
        else:     # a reference-style list (or 2-D array) over ALL trials: keep this rank's block
            lo_ = trials.offset if len(prevOptimRes) == trials.R_total else 0
            if isinstance(prevOptimRes, np.ndarray) and prevOptimRes.ndim == 2:
                x0 = _f64(prevOptimRes[lo_:lo_ + trials.R].reshape(trials.R, p.q, T))
            else:
                x0 = _f64(np.stack([np.asarray(prevOptimRes[i], dtype=np.float64).reshape(p.q, T)
                                    for i in range(lo_, lo_ + trials.R)]))
    # the per-trial T x T blocks of post_vsmGP are produced on first access (infRes['post_vsmGP']); the M-step only needs
    # their trial-sum, which the E-step delivers directly
    est = trials.estep_laplace(p, x0=x0, tol=tol, want_vsmGP=False, want_pautosum=True)
    if verbose:
        print('laplace inference: %d trials, Newton iterations max %d, factorisations %d'
              % (trials.R, est.stats['max_newton_iters'], est.stats['factorizations']))
    infRes = InfRes(est)
    post_lik = trials.post_lik(est)
    if returnOptimRes:
        return infRes, post_lik, _TrialView(est.x, lambda a: a.reshape(-1))
    return infRes, post_lik


# ----------------------------------------------------------------------------------------------
# Dual variational inference
# ----------------------------------------------------------------------------------------------
def _infer_T(C_big):
    """T from the Kronecker structure C_big = kron(C, I_T)^T (funs/util.py:595)."""
    C_big = np.asarray(C_big)
    rows, cols = C_big.shape
    g = np.gcd(rows, cols)
    for T in sorted({t for t in range(1, g + 1) if g % t == 0}, reverse=True):
        q, N = rows // T, cols // T
        blk = C_big.reshape(q, T, N, T)
        diag = np.einsum('ktnt->knt', blk)
        if np.count_nonzero(blk) == np.count_nonzero(diag) and np.all(diag == diag[:, :, :1]):
            return T
    raise ValueError("C_big is not of the form kron(C, eye(T)).T")


def _unpack_vi(C_big, K_big=None, K_bigInv=None, d_big=None):
    T = _infer_T(C_big)
    C_big = np.asarray(C_big)
    q, N = C_big.shape[0] // T, C_big.shape[1] // T
    C = C_big[::T, ::T].T.copy()
    blocks = lambda M: np.stack([np.asarray(M)[k * T:(k + 1) * T, k * T:(k + 1) * T] for k in range(q)])
    K = blocks(K_big) if K_big is not None else None
    Kinv = blocks(K_bigInv) if K_bigInv is not None else None
    d = np.asarray(d_big)[::T].copy() if d_big is not None else np.zeros(N)
    return C, d, K, Kinv, q, N, T


def VIPostCov(K_bigInv, C_big, lamb):
    """(postCovariance, postPrecision) — funs/inference.py:188-191 (relative 1e-6 diagonal jitter before inverting)."""
    C, d, _, Kinv, q, N, T = _unpack_vi(C_big, K_bigInv=K_bigInv)
    lam = _f64(np.asarray(lamb, dtype=np.float64).reshape(1, N, T))
    Kd = _f64(Kinv)
    K = kn.spd_inverse(Kd)[0]
    _, _, _, _, cov = kn.dualvi_eval(lam, torch.zeros_like(lam), _f64(C), _f64(d), K, Kd, want_grad=False, want_cov=True)
    return cov[0].cpu().numpy(), kn.hessian_dense(Kd, kn.rate_blocks(lam, _f64(C)))[0].cpu().numpy()


def VIPostMean(K_big, C_big, y_bar, lamb):
    """-K_big C_big (lamb - y)  — funs/inference.py:193-194."""
    C, d, K, _, q, N, T = _unpack_vi(C_big, K_big=K_big)
    lam = _f64(np.asarray(lamb, dtype=np.float64).reshape(1, N, T))
    y = _f64(np.asarray(y_bar, dtype=np.float64).reshape(1, N, T))
    Kd = _f64(K)
    Kinv = kn.spd_inverse(Kd)[0]
    _, _, mean, _, _ = kn.dualvi_eval(lam, y, _f64(C), _f64(d), Kd, Kinv, want_grad=False)
    return mean.reshape(-1).cpu().numpy()


def _dual_eval(lamb, ybar, C_big, K_big, K_bigInv, d_big):
    C, d, K, Kinv, q, N, T = _unpack_vi(C_big, K_big, K_bigInv, d_big)
    lam = _f64(np.asarray(lamb, dtype=np.float64).reshape(1, N, T))
    y = _f64(np.asarray(ybar, dtype=np.float64).reshape(1, N, T))
    D, grad, _, _, _ = kn.dualvi_eval(lam, y, _f64(C), _f64(d), _f64(K), _f64(Kinv))
    return float(D[0]), grad.reshape(-1).cpu().numpy()


def dualProblem(lamb, ybar, C_big, K_big, K_bigInv, d_big):
    """funs/inference.py:196-213."""
    return _dual_eval(lamb, ybar, C_big, K_big, K_bigInv, d_big)[0]


def dualProblem_grad(lamb, ybar, C_big, K_big, K_bigInv, d_big):
    """funs/inference.py:215-219."""
    return _dual_eval(lamb, ybar, C_big, K_big, K_bigInv, d_big)[1]


def dualProblemRho(rho, ybar, C_big, K_big, K_bigInv, d_big):
    """funs/inference.py:222-244."""
    return _dual_eval(np.exp(rho), ybar, C_big, K_big, K_bigInv, d_big)[0]


def dualProblemRho_grad(rho, ybar, C_big, K_big, K_bigInv, d_big):
    """funs/inference.py:246-256."""
    return _dual_eval(np.exp(rho), ybar, C_big, K_big, K_bigInv, d_big)[1] * np.exp(rho)


def dualVariational(experiment, params, optimizeLogLambda=False, prevOptimRes=None, returnOptimRes=True, verbose=False,
                    tol=1e-10, reducer=None):
    """varInfRes, -post_lik, var_lowerBound[, varOptimRes] — funs/inference.py:259-432.

    The reference runs L-BFGS-B on the dual per trial (bounded in lambda, or unbounded in rho = log lambda when
    ``optimizeLogLambda``); both variants have the same unique optimum, which is computed here directly as the
    stationary point of the dual (see csrc/dualvi.cu).  ``varOptimRes`` holds lambda* (or rho*) per trial."""
    trials = device_trials(experiment, reducer)
    T = trials.T
    p = device_params(params, T, experiment.binSize)
    params['tau'] = np.ndarray.flatten(np.asarray(params['tau'], dtype=np.float64))
    lam0 = None
    if prevOptimRes is not None:
        if isinstance(prevOptimRes, _TrialView) and prevOptimRes.tensor.shape[0] == trials.R:
            lam0 = prevOptimRes.tensor.reshape(trials.R, trials.N, T)
        else:
            sel = range(trials.offset, trials.offset + trials.R) if len(prevOptimRes) == trials.R_total else range(trials.R)
            lam0 = _f64(np.stack([np.asarray(prevOptimRes[i], dtype=np.float64).reshape(trials.N, T) for i in sel]))
        if optimizeLogLambda:
            lam0 = kn.emap("exp", lam0.contiguous())
        lam0 = lam0.contiguous()
    est = trials.estep_variational(p, lam0=lam0, tol=tol)
    if verbose:
        print('dual variational inference: %d trials, %d sweeps' % (trials.R, est.stats['sweeps']))

    infRes = InfRes(est, diag_scale=1.0 + 1e-6, W_fn=lambda r: kn.rate_blocks(est.lam[r:r + 1].contiguous(), p.C))
    post_lik = trials.post_lik(est)
    lower = trials.reducer.sum_scalar(float(est.dual.sum())) / trials.R_total
    if returnOptimRes:
        opt = kn.emap("log", est.lam) if optimizeLogLambda else est.lam
        return infRes, post_lik, lower, _TrialView(opt.reshape(trials.R, -1))
    return infRes, post_lik, lower
