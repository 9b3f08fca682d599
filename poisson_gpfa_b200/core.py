"""Device-resident EM steps (host orchestration only; all arithmetic is in libpgpfa_b200.so).

``DeviceTrials`` owns one rank's shard of the spike counts in HBM.  ``estep_laplace`` runs the batched
Newton E-step (funs/inference.py:67-185), ``mstep_cd`` the per-neuron Newton on C,d
(funs/learning.py:93-141 / :536-676 'useDiag'), ``mstep_tau`` the timescale update
(funs/learning.py:257-293 / :771-830).  Cross-rank reductions go through ``dist.Reducer``.
"""
import contextlib
import ctypes
import math
import os

import numpy as np
import torch

from . import _lib, kernels as kn
from ._lib import call, empty, handle, ptr, stream
from .dist import Reducer

EPS_NOISE = 0.001   # funs/util.py:599, funs/learning.py:286
LOWRANK_DELTA = 1e-14   # residual of the pivoted Cholesky of the smooth part of K (relative to its unit diagonal)


_side_streams = {}


def _side_stream():
    """One high-priority stream per device for the whole process (mini-batch EM creates many DeviceTrials; a stream
    per object would leave a trail of per-stream allocator pools)."""
    dev = torch.cuda.current_device()
    if dev not in _side_streams:
        _side_streams[dev] = torch.cuda.Stream(priority=-1)
    return _side_streams[dev]


class DeviceParams:
    """C (N,q), d (N), tau (q, seconds) on the device plus the derived prior blocks."""

    def __init__(self, C, d, tau, T, binSize):
        self.C = _lib.dev_f64(C)
        self.d = _lib.dev_f64(np.ravel(d) if not isinstance(d, torch.Tensor) else d.reshape(-1))
        self.tau = _lib.dev_f64(np.ravel(tau) if not isinstance(tau, torch.Tensor) else tau.reshape(-1))
        self.T, self.binSize = int(T), float(binSize)
        self._K = self._Kinv = self._logdetK = None
        self._lowrank = False

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

    @property
    def lowrank(self):
        """(F, Ft, ranks, eps) with K_k - eps I = F_k F_k^T (pivoted Cholesky, residual <= 1e-14), or None when the
        prior's numerical rank is not small (sum of ranks > qT/2: short timescales) and the dense path is the cheaper
        one.  PGPFA_LOWRANK=0 disables it."""
        if self._lowrank is False:
            self._lowrank = None
            if os.environ.get("PGPFA_LOWRANK", "1") != "0":
                F, Ft, ranks = kn.prior_lowrank(self.K, EPS_NOISE, LOWRANK_DELTA)
                if 0 < sum(ranks) <= (self.q * self.T) // 2:
                    self._lowrank = (F, Ft, ranks, EPS_NOISE)
        return self._lowrank

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
        self.side = None      # stream on which x / f / vsm are complete while vsmGP may still be in flight on the main one

    def means_stream(self):
        """Context in which work that needs only x, f and vsm runs: the side stream if the E-step left one."""
        return torch.cuda.stream(self.side) if self.side is not None else contextlib.nullcontext()


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
        self._cd_ws = None
        self._tau_ws = None

    def _means_stream(self):
        """High-priority stream ordered after the point of the last Laplace solve where the posterior means and
        time-diagonal covariances are final (pgpfa_stream_wait_means): the info check, the objective and the C,d
        M-step run there, underneath the selected-inverse kernel that is still producing post_vsmGP."""
        if os.environ.get("PGPFA_SIDE_STREAM", "1") == "0":      # debugging switch: everything on the caller's stream
            return None
        side = _side_stream()
        call("pgpfa_stream_wait_means", handle(), side.cuda_stream)
        return side

    # ------------------------------------------------------------------ E-step
    def estep_laplace(self, params, x0=None, tol=1e-8, max_newton=60, want_vsmGP=True, inexact_newton=True):
        R, N, T = self.y.shape
        q = params.q
        if self._lap_ws is None or self._lap_ws[0] != (R, q, T):
            full = _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, R)
            free, _ = torch.cuda.mem_get_info()
            budget = int(free * 0.80) - R * q * T * T * 8      # leave room for vsmGP
            nbytes = full if full <= budget else max(budget, _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, 1))
            self._lap_ws = ((R, q, T), _lib.workspace(nbytes))
        res = kn.laplace_solve(self.y, params.C, params.d, params.Kinv, x0=x0, tol=tol, max_newton=max_newton,
                               want_vsm=True, want_vsmGP=want_vsmGP, ws=self._lap_ws[1], inexact_newton=inexact_newton,
                               lowrank=params.lowrank)
        est = EStepResult(res.x, res.f, res.vsm, res.vsmGP, res.niter, res.stats, params, self)
        est.side = self._means_stream()
        with est.means_stream():
            if int(res.info.abs().max()) != 0:
                raise FloatingPointError("posterior Hessian not positive definite for %d trial(s)"
                                         % int((res.info != 0).sum()))
        return est

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
        with est.means_stream():
            return -self.reducer.sum_scalar(float(est.f.sum())) / self.R_total

    # ------------------------------------------------------------------ M-step C,d
    def mstep_cd(self, params, est, prior_w=0.0, tol=1e-10, max_iter=100, one_step=False, step_size=1.0,
                 prior_mat=None):
        """Per-neuron damped Newton on MStepObservationCost (+ 0.5*prior_w*|theta-theta_old|^2).
        Returns (C, d, cost, iterations).  `one_step`: a single (scaled) Newton step from the old
        parameters, the 'grad' online rule of funs/learning.py:884-891 with the analytic Hessian."""
        main = torch.cuda.current_stream()
        with est.means_stream():
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
                    th_cur = kn.emap("axpy", theta0, step, step_size)
                    break
                call("pgpfa_mstep_cd_update", ptr(stats), inv_R, float(prior_w), ptr(prior_mat), ptr(theta0), ptr(th_cur),
                     ptr(th_try), ptr(fcur), ptr(step), ptr(alpha), ptr(slope), ptr(done), 1 if it == 1 else 0, float(tol),
                     N, q, ptr(n_open), stream())
                if int(n_open.item()) == 0:
                    break
            cost = float(fcur.sum())
            C_new, d_new = th_cur[:, :q].contiguous(), th_cur[:, q].contiguous()
        if est.side is not None:                 # results were produced on the side stream: order the main one after it
            main.wait_stream(est.side)
            for t in (C_new, d_new) + ((hess,) if hess is not None else ()):
                t.record_stream(main)
        return C_new, d_new, cost, it, hess

    def cd_cost_grad(self, theta, est):
        """(cost, grad (N,q+1)) of MStepObservationCost at theta, normalised by the global trial count."""
        q = theta.shape[1] - 1
        stats = self.reducer.sum_tensor(kn.mstep_cd_stats(self.y, est.x, est.vsm, theta))
        return float(stats[0].sum()) / self.R_total, (stats[1:q + 2].T / self.R_total).contiguous(), stats

    # ------------------------------------------------------------------ M-step tau
    def pautosum(self, est):
        P = kn.pautosum(est.vsmGP, est.x)
        return self.reducer.sum_tensor(P)

    def mstep_tau(self, params, Psum, numTrials=None, prior_step=None, xtol=1e-10, max_rounds=14, ncand=9):
        """q independent scalar minimisations over p = log(1/tau_bins^2) (funs/learning.py:257-293, :771-830).
        The reference hands each to scipy (BFGS / TNC) from p0 = the old tau; the result is the first zero of
        the gradient in the descent direction from p0.  Here all latents advance in lock-step and every device
        launch evaluates `ncand` candidate points per latent (the T x T factorisations are latency-bound, so
        candidates are free): round 1 brackets the sign change of the gradient around p0, later rounds place
        the candidates around the inverse-cubic interpolant of the bracket's neighbours (error ~ width^4), so
        3-4 launches reach |dp| < 2e-11.  Returns (tau_seconds (q), details)."""
        q, T = params.q, self.T
        R = float(self.R_total if numTrials is None else numTrials)
        m = int(ncand)
        key = (q, T, m)
        if self._tau_ws is None or self._tau_ws[0] != key:
            self._tau_ws = (key, _lib.workspace(_lib.lib.pgpfa_tau_eval_workspace_bytes(q * m, T)))
        tau_old = params.tau
        P_rep = Psum.repeat(m, 1, 1).contiguous()               # slot = c*q + k  (data movement only)
        tau_old_rep = tau_old.repeat(m).contiguous()
        pw = 0.0 if prior_step is None else 1.0 / float(prior_step) ** 2
        nev = [0]

        def fg(cands):                                          # (m,q) -> f, g (m,q)
            nev[0] += 1
            c, g = kn.tau_eval(_lib.dev_f64(np.ascontiguousarray(cands).reshape(-1)), P_rep, R, T, EPS_NOISE, pw,
                               tau_old_rep, self.binSize, ws=self._tau_ws[1])
            both = torch.stack([c, g]).cpu().numpy()
            return both[0].reshape(m, q), both[1].reshape(m, q)

        oldTau_bins = tau_old.cpu().numpy() * 1000.0 / self.binSize
        p0 = np.log(1.0 / oldTau_bins ** 2)
        offs = np.array([0.0, -0.1, 0.1, -0.25, 0.25, -0.5, 0.5, -1.0, 1.0])[:m]
        cands = p0[None, :] + offs[:, None]
        f, g = fg(cands)
        pts = [sorted(zip(cands[:, k], g[:, k], f[:, k])) for k in range(q)]       # per latent: (p, g, f) ascending in p
        g0, f0 = g[0].copy(), f[0].copy()
        done = np.zeros(q, dtype=bool)
        p_star = p0.copy()
        bracketed = np.zeros(q, dtype=bool)

        def bracket_of(k):
            """Index i with a sign change of g between pts[k][i] and pts[k][i+1], nearest to p0 on the descent side."""
            P = pts[k]
            idx0 = min(range(len(P)), key=lambda i: abs(P[i][0] - p0[k]))
            if P[idx0][1] == 0.0:
                return ('exact', idx0)
            rng = range(idx0, len(P) - 1) if P[idx0][1] < 0 else range(idx0 - 1, -1, -1)
            for i in rng:
                if P[i][1] == 0.0:
                    return ('exact', i)
                if P[i][1] < 0.0 <= P[i + 1][1]:
                    return ('br', i)
            return ('none', len(P) - 1 if P[idx0][1] < 0 else 0)

        def interpolate(k, i):
            """Zero of g inside (P[i], P[i+1]) by inverse polynomial interpolation through up to 4 neighbours."""
            P = pts[k]
            a, ga = P[i][0], P[i][1]
            b, gb = P[i + 1][0], P[i + 1][1]
            sel = P[max(0, i - 1):i + 3]
            gs = np.array([t[1] for t in sel]); ps = np.array([t[0] for t in sel])
            c = a - ga * (b - a) / (gb - ga)
            err = 0.5 * (b - a) ** 2                              # secant: error ~ |g''/2g'| (c-a)(b-c)
            if len(sel) >= 3 and np.all(np.diff(gs) > 0):
                est = 0.0
                for u in range(len(sel)):                       # Lagrange form of p(g) at g = 0
                    wgt = 1.0
                    for v in range(len(sel)):
                        if v != u:
                            wgt *= (0.0 - gs[v]) / (gs[u] - gs[v])
                    est += wgt * ps[u]
                if a < est < b:
                    c = est
                    span = ps.max() - ps.min()
                    err = 0.25 * (b - a) ** 2 * span ** (len(sel) - 2)   # ~ product of the distances to the nodes
            return c, a, b, err

        for rnd in range(max_rounds):
            cands = np.tile(p_star[None, :], (m, 1))
            for k in range(q):
                if done[k]:
                    continue
                kind, i = bracket_of(k)
                P = pts[k]
                if kind == 'exact':
                    p_star[k], done[k], bracketed[k] = P[i][0], True, True
                    continue
                if kind == 'none':                               # walk further downhill with growing steps
                    edge = P[i][0]
                    span = max(0.5, abs(edge - p0[k]))
                    sgn = 1.0 if P[i][1] < 0 else -1.0
                    cands[:, k] = np.clip(edge + sgn * span * (0.5 * 1.7 ** np.arange(m)), -40.0, 20.0)
                    if abs(edge) >= 20.0:
                        done[k] = True                           # monotone cost: keep the old tau (flagged)
                    continue
                bracketed[k] = True
                c, a, b, err = interpolate(k, i)
                w = b - a
                p_star[k] = c
                if err <= xtol * (1.0 + abs(c)) or w <= xtol * (1.0 + abs(a)):
                    done[k] = True
                    continue
                h1 = min(max(2.0 * err, 4.0 * xtol * (1.0 + abs(c))), w / 16.0)
                h2 = min(max(4.0 * h1, 0.25 * w ** 2), w / 4.0)
                hs = [h1, h2] + [min(h2 * 4.0 ** e, w / 2.5) for e in range(1, (m - 1) // 2 - 1)]
                pr = np.array([c] + [c + sg * hh for hh in hs for sg in (-1.0, 1.0)])[:m]
                lo, hi = a + 1e-3 * w, b - 1e-3 * w
                cands[:, k] = np.clip(pr, lo, hi)
            if done.all():
                break
            f, g = fg(cands)
            for k in range(q):
                if not done[k]:
                    have = {t[0] for t in pts[k]}
                    pts[k] = sorted(pts[k] + [(cands[c_, k], g[c_, k], f[c_, k]) for c_ in range(m) if cands[c_, k] not in have])
        p_new = np.where(bracketed, p_star, p0)
        tau_bins = (1.0 / np.exp(p_new)) ** 0.5
        fun = np.array([min(pts[k], key=lambda t: abs(t[0] - p_new[k]))[2] for k in range(q)])
        gr = np.array([min(pts[k], key=lambda t: abs(t[0] - p_new[k]))[1] for k in range(q)])
        details = {'p': p_new, 'p0': p0, 'grad': gr, 'fun': fun, 'fun0': f0, 'grad0': g0, 'nfev': nev[0],
                   'bracketed': bracketed}
        return tau_bins * self.binSize / 1000.0, details
