"""Device-resident EM steps (host orchestration only; all arithmetic is in libpgpfa_b200.so).

``DeviceTrials`` owns one rank's shard of the spike counts in HBM.  ``estep_laplace`` runs the batched
Newton E-step (funs/inference.py:67-185), ``mstep_cd`` the per-neuron Newton on C,d
(funs/learning.py:93-141 / :536-676 'useDiag'), ``mstep_tau`` the timescale update
(funs/learning.py:257-293 / :771-830).  Cross-rank reductions go through ``dist.Reducer``.
"""
import ctypes
import math

import numpy as np
import torch

from . import _lib, kernels as kn
from ._lib import call, empty, ptr, stream
from .dist import Reducer

EPS_NOISE = 0.001   # funs/util.py:599, funs/learning.py:286


class DeviceParams:
    """C (N,q), d (N), tau (q, seconds) on the device plus the derived prior blocks."""

    def __init__(self, C, d, tau, T, binSize):
        self.C = _lib.dev_f64(C)
        self.d = _lib.dev_f64(np.ravel(d) if not isinstance(d, torch.Tensor) else d.reshape(-1))
        self.tau = _lib.dev_f64(np.ravel(tau) if not isinstance(tau, torch.Tensor) else tau.reshape(-1))
        self.T, self.binSize = int(T), float(binSize)
        self._K = self._Kinv = self._logdetK = None

    @property
    def q(self):
        return self.C.shape[1]

    @property
    def K(self):
        if self._K is None:
            self._K = kn.make_K(self.tau, self.T, self.binSize, EPS_NOISE)
        return self._K

    @property
    def Kinv(self):
        """K^-1 per latent: q independent T x T Cholesky inverses (the reference inverts the block-diagonal
        qT x qT K_big with one dense LU, funs/inference.py:82)."""
        if self._Kinv is None:
            self._Kinv, self._logdetK, info = kn.spd_inverse(self.K)
            if int(info.abs().max()) != 0:
                raise FloatingPointError("GP prior covariance K is not positive definite (tau=%s)" % self.tau.tolist())
        return self._Kinv

    def to_numpy_dict(self):
        return {'C': self.C.cpu().numpy(), 'd': self.d.cpu().numpy(), 'tau': self.tau.cpu().numpy()}

    @property
    def theta(self):
        return torch.cat([self.C, self.d[:, None]], dim=1).contiguous()


class EStepResult:
    """Posterior statistics of one E-step, device resident (layouts of include/pgpfa_b200.h)."""

    def __init__(self, x, f, vsm, vsmGP, niter, stats, params, trials):
        self.x, self.f, self.vsm, self.vsmGP = x, f, vsm, vsmGP
        self.niter, self.stats = niter, stats
        self.params, self.trials = params, trials


class DeviceTrials:
    """One rank's trials: y (R_local, N, T) float64 in HBM (counts are promoted to float64 exactly as the
    reference does on use; s_y = 8 bytes in the roofline byte counts)."""

    def __init__(self, y, binSize, reducer=None, R_total=None, offset=0):
        self.y = y if isinstance(y, torch.Tensor) else _lib.dev_f64(y)
        assert self.y.dim() == 3 and self.y.is_contiguous()
        self.R, self.N, self.T = self.y.shape
        self.binSize = float(binSize)
        self.reducer = reducer if reducer is not None else Reducer()
        self.offset = int(offset)           # index of this shard's first trial in the full experiment
        self.R_total = int(R_total) if R_total is not None else int(self.reducer.sum_scalar(self.R))
        self._lap_ws = None
        self._factor_key = None      # (R,q,T) when the Laplace workspace holds every trial's factor at its last mode
        self._cd_ws = None
        self._tau_ws = None

    # ------------------------------------------------------------------ E-step
    def estep_laplace(self, params, x0=None, tol=1e-8, max_newton=60, want_vsmGP=True, reuse_factor=True):
        R, N, T = self.y.shape
        q = params.q
        if self._lap_ws is None or self._lap_ws[0] != (R, q, T):
            full = _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, R)
            free, _ = torch.cuda.mem_get_info()
            budget = int(free * 0.80) - R * q * T * T * 8      # leave room for vsmGP
            nbytes = full if full <= budget else max(budget, _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, 1))
            self._lap_ws = ((R, q, T), _lib.workspace(nbytes))
            self._factor_key = None
        # factors kept from the previous E-step of the SAME trials drive cheap chord iterations (warm start only)
        reuse = reuse_factor and x0 is not None and self._factor_key == (R, q, T)
        self._factor_key = None
        res = kn.laplace_solve(self.y, params.C, params.d, params.Kinv, x0=x0, tol=tol, max_newton=max_newton,
                               want_vsm=True, want_vsmGP=want_vsmGP, ws=self._lap_ws[1], reuse_factor=reuse)
        if res.rc == 0 and res.stats["factors_kept"]:
            self._factor_key = (R, q, T)
        if int(res.info.abs().max()) != 0:
            raise FloatingPointError("posterior Hessian not positive definite for %d trial(s)"
                                     % int((res.info != 0).sum()))
        return EStepResult(res.x, res.f, res.vsm, res.vsmGP, res.niter, res.stats, params, self)

    def estep_variational(self, params, lam0=None, tol=1e-10, max_iter=300, want_vsmGP=True):
        """Dual variational E-step (funs/inference.py:259-432): the stationary point of the dual for every
        trial of the shard.  Returns (EStepResult with the VARIATIONAL mean/covariance slices, lam (R,N,T),
        dual values (R))."""
        res = kn.dualvi_solve(self.y, params.C, params.d, params.K, params.Kinv, lam0=lam0, tol=tol,
                              max_iter=max_iter, want_vsmGP=want_vsmGP)
        if int(res.info.abs().max()) != 0:
            raise FloatingPointError("variational posterior precision not positive definite")
        if res.rc != 0:
            raise RuntimeError("dual variational fixed point: %d trial(s) not converged in %d sweeps"
                               % (res.stats["not_converged"], max_iter))
        est = EStepResult(res.mean, res.f, res.vsm, res.vsmGP, res.niter, res.stats, params, self)
        est.lam, est.dual = res.lam, res.D
        return est

    def post_lik(self, est):
        """-mean_r L(x_r*) over ALL trials (funs/inference.py:175,183)."""
        return -self.reducer.sum_scalar(float(est.f.sum())) / self.R_total

    # ------------------------------------------------------------------ M-step C,d
    def mstep_cd(self, params, est, prior_w=0.0, tol=1e-10, max_iter=100, one_step=False, step_size=1.0,
                 prior_mat=None):
        """Per-neuron damped Newton on MStepObservationCost (+ 0.5*prior_w*|theta-theta_old|^2).
        Returns (C, d, cost, iterations).  `one_step`: a single (scaled) Newton step from the old
        parameters, the 'grad' online rule of funs/learning.py:884-891 with the analytic Hessian."""
        N, q = params.C.shape
        P = q + 1
        theta0 = params.theta
        th_cur, th_try = theta0.clone(), theta0.clone()
        fcur, alpha, slope = empty(N), empty(N), empty(N)
        step = empty(N, P)
        done = torch.zeros(N, dtype=torch.int32, device="cuda")
        n_open = torch.zeros(1, dtype=torch.int32, device="cuda")
        if self._cd_ws is None or self._cd_ws[0] != (q, N):
            self._cd_ws = ((q, N), _lib.workspace(_lib.lib.pgpfa_mstep_cd_workspace_bytes(q, N)))
        inv_R = 1.0 / self.R_total
        it = 0
        hess = None
        for it in range(1, max_iter + 1):
            stats = kn.mstep_cd_stats(self.y, est.x, est.vsm, th_try, ws=self._cd_ws[1])
            stats = self.reducer.sum_tensor(stats)
            if one_step:
                hess = stats
                call("pgpfa_mstep_cd_update", ptr(stats), inv_R, float(prior_w), ptr(prior_mat), ptr(theta0), ptr(th_cur),
                     ptr(th_try), ptr(fcur), ptr(step), ptr(alpha), ptr(slope), ptr(done), 1, 0.0, N, q, ptr(n_open),
                     stream())
                th_cur = theta0 + step_size * step
                break
            call("pgpfa_mstep_cd_update", ptr(stats), inv_R, float(prior_w), ptr(prior_mat), ptr(theta0), ptr(th_cur),
                 ptr(th_try), ptr(fcur), ptr(step), ptr(alpha), ptr(slope), ptr(done), 1 if it == 1 else 0, float(tol),
                 N, q, ptr(n_open), stream())
            if int(n_open.item()) == 0:
                break
        cost = float(fcur.sum())
        return th_cur[:, :q].contiguous(), th_cur[:, q].contiguous(), cost, it, hess

    def cd_cost_grad(self, theta, est):
        """(cost, grad (N,q+1)) of MStepObservationCost at theta, normalised by the global trial count."""
        q = theta.shape[1] - 1
        stats = self.reducer.sum_tensor(kn.mstep_cd_stats(self.y, est.x, est.vsm, theta))
        return float(stats[0].sum()) / self.R_total, (stats[1:q + 2].T / self.R_total).contiguous(), stats

    # ------------------------------------------------------------------ M-step tau
    def pautosum(self, est):
        P = kn.pautosum(est.vsmGP, est.x)
        return self.reducer.sum_tensor(P)

    def mstep_tau(self, params, Psum, numTrials=None, prior_step=None, gtol=1e-9, max_eval=80):
        """q independent scalar minimisations over p = log(1/tau_bins^2) (funs/learning.py:257-293,
        :771-830), all latents in lock-step: bracket the first zero of the gradient in the descent
        direction from the old tau, then Illinois (safeguarded secant) refinement.  One device
        evaluation per iteration serves all latents.  Returns (tau_seconds (q), details)."""
        q, T = params.q, self.T
        R = float(self.R_total if numTrials is None else numTrials)
        if self._tau_ws is None or self._tau_ws[0] != (q, T):
            self._tau_ws = ((q, T), _lib.workspace(_lib.lib.pgpfa_tau_eval_workspace_bytes(q, T)))
        tau_old = params.tau
        pw = 0.0 if prior_step is None else 1.0 / float(prior_step) ** 2
        nev = [0]

        def fg(p_np):
            nev[0] += 1
            c, g = kn.tau_eval(_lib.dev_f64(p_np), Psum, R, T, EPS_NOISE, pw, tau_old, self.binSize,
                               ws=self._tau_ws[1])
            return c.cpu().numpy(), g.cpu().numpy()

        oldTau_bins = tau_old.cpu().numpy() * 1000.0 / self.binSize
        p0 = np.log(1.0 / oldTau_bins ** 2)
        p, (f, g) = p0.copy(), fg(p0)
        f0 = f.copy()
        # bracket: walk downhill with doubling steps until the gradient changes sign
        a, ga = p.copy(), g.copy()
        b, gb = p.copy(), g.copy()
        stepsz = np.full(q, 0.25)
        have = np.abs(g) <= 0.0
        for _ in range(30):
            if have.all():
                break
            trial = np.where(have, b, b - np.sign(ga) * stepsz)
            trial = np.clip(trial, -40.0, 20.0)
            ft, gt = fg(trial)
            flip = (np.sign(gt) != np.sign(ga)) | (gt == 0.0)
            upd = ~have
            # keep `a` as the last point with the original sign, `b` the first point with the other sign
            mv = upd & ~flip
            a[mv], ga[mv] = trial[mv], gt[mv]
            b[upd], gb[upd] = trial[upd], gt[upd]
            have = have | flip
            stepsz = np.where(have, stepsz, stepsz * 2.0)
            if nev[0] >= max_eval:
                break
        # Illinois (safeguarded secant) iterations on g over [a,b], sign(ga) != sign(gb); a latent stops
        # when the proposed move is below xtol (superlinear convergence: the move estimates the error)
        side = np.zeros(q, dtype=int)
        pick_a = np.abs(ga) < np.abs(gb)
        x, gx, fx = np.where(pick_a, a, b), np.where(pick_a, ga, gb), f.copy()
        conv = ~have | (gx == 0.0)
        xtol = 1e-11
        while not conv.all() and nev[0] < max_eval:
            denom = gb - ga
            with np.errstate(divide='ignore', invalid='ignore'):
                c = np.where(denom != 0.0, b - gb * (b - a) / denom, 0.5 * (a + b))
            lo, hi = np.minimum(a, b), np.maximum(a, b)
            c = np.where((c > lo) & (c < hi), c, 0.5 * (a + b))
            small = (np.abs(c - x) <= xtol * (1.0 + np.abs(x))) | (np.abs(b - a) <= xtol * (1.0 + np.abs(a)))
            x = np.where(~conv & small, c, x)
            conv = conv | small
            if conv.all():
                break
            c = np.where(conv, x, c)
            fc, gc = fg(c)
            for k in range(q):
                if conv[k]:
                    continue
                if gc[k] == 0.0:
                    conv[k] = True
                elif np.sign(gc[k]) == np.sign(gb[k]):
                    b[k], gb[k] = c[k], gc[k]
                    if side[k] == -1:
                        ga[k] *= 0.5
                    side[k] = -1
                else:
                    a[k], ga[k] = c[k], gc[k]
                    if side[k] == 1:
                        gb[k] *= 0.5
                    side[k] = 1
                x[k], gx[k], fx[k] = c[k], gc[k], fc[k]
        p_new = np.where(have, x, p0)
        tau_bins = (1.0 / np.exp(p_new)) ** 0.5
        details = {'p': p_new, 'p0': p0, 'grad': gx, 'fun': fx, 'fun0': f0, 'nfev': nev[0], 'bracketed': have}
        return tau_bins * self.binSize / 1000.0, details
