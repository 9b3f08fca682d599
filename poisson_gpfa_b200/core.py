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


def read_packed(tensors):
    """ONE device -> host copy (one synchronisation) for a list of small device tensors; returns float64 numpy arrays
    of the original shapes.  Integer tensors are converted (exact below 2^53)."""
    flat = torch.cat([t.detach().reshape(-1).to(torch.float64) for t in tensors])
    host = _lib.to_host(flat)
    out, o = [], 0
    for t in tensors:
        n = t.numel()
        out.append(host[o:o + n].reshape(tuple(t.shape)))
        o += n
    return out


_PRIOR_SIDE = {}


def _prior_side_stream():
    dev = torch.cuda.current_device()
    if dev not in _PRIOR_SIDE:
        _PRIOR_SIDE[dev] = torch.cuda.Stream(device=dev)
    return _PRIOR_SIDE[dev]


class DeviceParams:
    """C (N,q), d (N), tau (q, seconds) on the device plus the derived prior blocks.

    ``prepare()`` only ENQUEUES the prior construction (K, K^-1 by q independent T x T Cholesky inverses — the
    reference inverts the block-diagonal qT x qT K_big with one dense LU, funs/inference.py:82 — and the pivoted-Cholesky
    low-rank factor); the two small results the host needs (is K positive definite, the ranks) are read either together
    with other flags (``pending()`` / ``resolve()``, used by the device-resident EM step) or lazily on first use."""

    def __init__(self, C, d, tau, T, binSize):
        self.C = _lib.dev_f64(C)
        self.d = _lib.dev_f64(np.ravel(d) if not isinstance(d, torch.Tensor) else d.reshape(-1))
        self.tau = _lib.dev_f64(np.ravel(tau) if not isinstance(tau, torch.Tensor) else tau.reshape(-1))
        self.T, self.binSize = int(T), float(binSize)
        self._K = self._Kinv = self._logdetK = self._Kinfo = None
        self._lr_dev = None          # (F, Ft, ranks on the device)
        self._resolved = False
        self._lowrank = None
        self._use_lr = os.environ.get("PGPFA_LOWRANK", "1") != "0"

    @property
    def q(self):
        return self.C.shape[1]

    def prepare(self):
        if self._K is None:
            self._K = kn.make_K(self.tau, self.T, self.binSize, EPS_NOISE)
            if self._use_lr:
                # the pivoted Cholesky (one CTA per latent, ~0.12 ms of pure latency) runs on a side stream beside the
                # inverse of K (one CTA per latent as well); both only read K, and the caller's stream waits for both
                q, T = self.q, self.T
                out = (kn.empty(q, T, T), kn.empty(q, T, T), kn.empty(q, dtype=torch.int32))
                main, side = torch.cuda.current_stream(), _prior_side_stream()
                ev = torch.cuda.Event()
                ev.record(main)
                side.wait_event(ev)
                with torch.cuda.stream(side):
                    self._lr_dev = kn.prior_lowrank_async(self._K, EPS_NOISE, LOWRANK_DELTA, out=out)
                    ev_lr = torch.cuda.Event()
                    ev_lr.record(side)
            self._Kinv, self._logdetK, self._Kinfo = kn.spd_inverse(self._K)
            if self._use_lr:
                torch.cuda.current_stream().wait_event(ev_lr)
        return self

    def pending(self):
        """Small device tensors whose host values ``resolve`` needs (one packed read for all of them)."""
        self.prepare()
        return [self._Kinfo] + ([self._lr_dev[2]] if self._lr_dev is not None else [])

    def resolve(self, host_vals=None):
        if self._resolved:
            return self
        if host_vals is None:
            host_vals = read_packed(self.pending())
        if np.abs(host_vals[0]).max() != 0:
            raise FloatingPointError("GP prior covariance K is not positive definite (tau=%s)" % self.tau.tolist())
        if self._lr_dev is not None:
            ranks = [int(v) for v in host_vals[1]]
            # the low-rank pass only when the prior's numerical rank is small (sum of ranks <= qT/2); short timescales
            # make the dense tiled path the cheaper one
            if 0 < sum(ranks) <= (self.q * self.T) // 2:
                self._lowrank = (self._lr_dev[0], self._lr_dev[1], ranks, EPS_NOISE)
        self._resolved = True
        return self

    @property
    def K(self):
        return self.prepare()._K

    @property
    def Kinv(self):
        return self.resolve()._Kinv

    @property
    def lowrank(self):
        """(F, Ft, ranks, eps) with K_k - eps I = F_k F_k^T (pivoted Cholesky, residual <= 1e-14), or None when the
        prior's numerical rank is not small.  PGPFA_LOWRANK=0 disables it."""
        return self.resolve()._lowrank

    def to_numpy_dict(self):
        C, d, tau = read_packed([self.C, self.d, self.tau])
        return {'C': C, 'd': d, 'tau': tau}

    @property
    def theta(self):
        return torch.cat([self.C, self.d[:, None]], dim=1).contiguous()


class EStepResult:
    """Posterior statistics of one E-step, device resident (layouts of include/pgpfa_b200.h)."""

    def __init__(self, x, f, vsm, vsmGP, niter, stats, params, trials, info=None):
        self.x, self.f, self.vsm, self.vsmGP = x, f, vsm, vsmGP
        self.niter, self.stats, self.info = niter, stats, info
        self.params, self.trials = params, trials
        self.side = None      # stream on which x / f / vsm are complete while vsmGP may still be in flight on the main one
        self._checked = False
        self.pautosum = None  # (q,T,T) local trial-sum of post_vsmGP + m m^T when the E-step produced it directly
        self.tol = 1e-8

    def get_vsmGP(self):
        """post_vsmGP (R,q,T,T).  The EM loop only consumes the trial-sum (PautoSum), which the low-rank pass produces
        directly; the per-trial T x T blocks are materialised on first access by one more posterior pass at the stored
        mode (a converged start: one evaluation, no further Newton step)."""
        if self.vsmGP is None and self.x is not None and self.x.shape[0] > 0:
            p, t = self.params, self.trials
            res = kn.laplace_solve(t.y, p.C, p.d, p.Kinv, x0=self.x, tol=self.tol, want_vsm=False, want_vsmGP=True,
                                   lowrank=p.lowrank)
            self.vsmGP = res.vsmGP
        elif self.vsmGP is None and self.x is not None:
            q, T = self.params.q, self.trials.T
            self.vsmGP = torch.zeros(0, q, T, T, dtype=torch.float64, device="cuda")
        return self.vsmGP

    def means_stream(self):
        """Context in which work that needs only x, f and vsm runs: the side stream if the E-step left one."""
        return torch.cuda.stream(self.side) if self.side is not None else contextlib.nullcontext()

    def flags(self):
        """Device tensor [sum_r f_r, 1 if a posterior Hessian of this shard was not positive definite, 1 if the Newton
        iteration limit was reached] (enqueued on the means stream).  It is summed over ranks together with the
        objective, so every rank sees a failure of any rank at the same read."""
        with self.means_stream():
            fl = torch.zeros(3, dtype=torch.float64, device="cuda")
            if self.f is not None and self.f.numel():
                fl[0] = self.f.sum()
                if self.info is not None:
                    fl[1] = (self.info != 0).any().to(torch.float64)
            if self.stats and self.stats.get("not_converged", 0) > 0:
                fl[2] = 1.0
        return fl

    @staticmethod
    def raise_on(flags_host, what="Laplace"):
        fsum, notpd, notconv = (float(v) for v in flags_host[:3])
        if notpd != 0:
            raise FloatingPointError("%s E-step: posterior Hessian / precision not positive definite for at least one trial"
                                     % what)
        if notconv != 0:
            raise RuntimeError("%s E-step: iteration limit reached before every trial converged" % what)
        if not np.isfinite(fsum):
            raise FloatingPointError("%s E-step: non-finite objective (rate overflow?)" % what)


class CdSolve:
    """State of a C,d Newton solve in flight on the device (DeviceTrials.mstep_cd_async)."""

    def __init__(self, trials, est, q, N):
        self.trials, self.est, self.q, self.N = trials, est, q, N
        self.issued = 0
        self.hess = None
        self.fsum = None

    def pending(self):
        """[n_open (int[4]), sum of the per-neuron costs, (objective sum, E-step error code)]"""
        with self.est.means_stream():
            out = [self.n_open, self.fcur.sum().reshape(1), self.extra]
        return out

    def finish(self, host_vals=None, max_iter=100):
        """Host side of the solve: read the flags (unless given), run more iterations while neurons are open (rare: the
        blind schedule covers the usual 3-4), hand back (C, d, cost, iterations, hess)."""
        t = self.trials
        if host_vals is None:
            pend = self.pending()          # enqueued on the solve's stream: join AFTER it, then read on the caller's
            self.join()
            host_vals = read_packed(pend)
        n_open, cost, extra = host_vals
        while n_open[0] > 0 and self.issued < max_iter:
            with self.est.means_stream():
                t._cd_iterate(self, 1)
            pend = self.pending()
            self.join()
            n_open, cost, extra = read_packed(pend)
        q = self.q
        with self.est.means_stream():
            C_new, d_new = self.th_cur[:, :q].contiguous(), self.th_cur[:, q].contiguous()
        self.join((C_new, d_new))
        self.extra_host = extra
        its = int(n_open[1]) if not self.one_step else 1
        return C_new, d_new, float(cost[0]), its, self.hess

    def join(self, tensors=()):
        """Order the caller's stream after the solve's (side) stream."""
        side = self.est.side
        if side is not None:
            self.main.wait_stream(side)
            for x in tensors:
                x.record_stream(self.main)


class TauSolve:
    """State of a timescale search in flight on the device (DeviceTrials.mstep_tau_async)."""

    def pending(self):
        return [self.flags]

    def finish(self, host_vals=None, max_rounds=14):
        t = self.trials
        if host_vals is None:
            host_vals = read_packed(self.pending())
        flags = host_vals[0]
        while flags[0] > 0 and self.rounds < max_rounds + 1:
            t._tau_rounds(self, 1)
            flags = read_packed(self.pending())[0]
        self.flags_host = flags
        return self.tau

    def details(self):
        """Host dictionary of the search (one read): p, p0, grad, fun, fun0, grad0, nfev, bracketed."""
        det, = read_packed([self.det])
        fl = self.flags_host
        q = det.shape[1]
        br = np.array([(int(fl[2]) >> k) & 1 for k in range(q)], dtype=bool)
        return {'p': det[0], 'p0': det[1], 'grad': det[2], 'fun': det[3], 'fun0': det[4], 'grad0': det[5],
                'nfev': int(fl[1]), 'bracketed': br}


class DeviceTrials:
    """One rank's trials: y (R_local, N, T) float64 in HBM (counts are promoted to float64 exactly as the
    reference does on use; s_y = 8 bytes in the roofline byte counts).  R_local may be 0 (a mini-batch smaller than
    the world): such a rank launches nothing, contributes zero statistics and still joins every collective."""

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
        self._tau_blind = 3       # evaluation rounds of the timescale search enqueued blind: what the last search needed

    def _means_stream(self):
        """High-priority stream ordered after the point of the last Laplace solve where the posterior means and
        time-diagonal covariances are final (pgpfa_stream_wait_means): the objective and the whole C,d M-step run
        there, underneath the batched GEMM that is still producing post_vsmGP."""
        if os.environ.get("PGPFA_SIDE_STREAM", "1") == "0":      # debugging switch: everything on the caller's stream
            return None
        side = _side_stream()
        call("pgpfa_stream_wait_means", handle(), side.cuda_stream)
        return side

    # ------------------------------------------------------------------ E-step
    def estep_laplace(self, params, x0=None, tol=1e-8, max_newton=60, want_vsmGP=True, inexact_newton=True,
                      want_pautosum=False):
        """Batched Laplace E-step.  `tol` bounds the last Newton step (relative, inf-norm); the returned mode is one
        polishing Newton step closer (~1e-13 measured), the covariances are those at the point before that step
        (within ~4e-11 of the converged ones at tol = 1e-8, tools/parity_report.py)."""
        R, N, T = self.y.shape
        q = params.q
        if R == 0:
            z = lambda *s: torch.zeros(*s, dtype=torch.float64, device="cuda")
            return EStepResult(z(0, q, T), z(0), z(0, T, q, q), z(0, q, T, T) if want_vsmGP else None,
                               torch.zeros(0, dtype=torch.int32, device="cuda"), {"not_converged": 0, "lowrank_r": 0}, params,
                               self, info=torch.zeros(0, dtype=torch.int32, device="cuda"))
        if self._lap_ws is None or self._lap_ws[0] != (R, q, T):
            full = _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, R)
            free, _ = torch.cuda.mem_get_info()
            budget = int(free * 0.80) - R * q * T * T * 8      # leave room for vsmGP
            nbytes = full if full <= budget else max(budget, _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, 1))
            self._lap_ws = ((R, q, T), _lib.workspace(nbytes))
        res = kn.laplace_solve(self.y, params.C, params.d, params.Kinv, x0=x0, tol=tol, max_newton=max_newton,
                               want_vsm=True, want_vsmGP=want_vsmGP, ws=self._lap_ws[1], inexact_newton=inexact_newton,
                               lowrank=params.lowrank, want_pautosum=want_pautosum)
        est = EStepResult(res.x, res.f, res.vsm, res.vsmGP if want_vsmGP else None, res.niter, res.stats, params, self,
                          info=res.info)
        est.pautosum, est.tol = res.pautosum, tol
        est.side = self._means_stream()
        return est

    def estep_variational(self, params, lam0=None, tol=1e-10, max_iter=300, want_vsmGP=True):
        """Dual variational E-step (funs/inference.py:259-432): the stationary point of the dual for every
        trial of the shard.  Returns (EStepResult with the VARIATIONAL mean/covariance slices, lam (R,N,T),
        dual values (R))."""
        if self.R == 0:
            est = self.estep_laplace(params)
            est.lam = torch.zeros(0, self.N, self.T, dtype=torch.float64, device="cuda")
            est.dual = torch.zeros(0, dtype=torch.float64, device="cuda")
            return est
        res = kn.dualvi_solve(self.y, params.C, params.d, params.K, params.Kinv, lam0=lam0, tol=tol,
                              max_iter=max_iter, want_vsmGP=want_vsmGP)
        res.stats["not_converged_sweeps"] = res.stats["not_converged"]
        est = EStepResult(res.mean, res.f, res.vsm, res.vsmGP, res.niter, res.stats, params, self, info=res.info)
        est.lam, est.dual = res.lam, res.D
        return est

    def post_lik(self, est, what="Laplace"):
        """-mean_r L(x_r*) over ALL trials (funs/inference.py:175,183).  One packed read: the objective sum travels with
        the E-step's error code through the same all-reduce, so a failure on any rank raises on every rank."""
        fl = est.flags()
        with est.means_stream():           # the read synchronises the stream the flags were produced on
            fl = self.reducer.sum_tensor(fl)
            fl = _lib.to_host(fl)
        EStepResult.raise_on(fl, what)
        est._checked = True
        return -float(fl[0]) / self.R_total

    # ------------------------------------------------------------------ M-step C,d
    def _cd_iterate(self, cd, n_iters):
        """Enqueue n_iters Newton iterations of a C,d solve (no host read)."""
        est = cd.est
        N, q = cd.N, cd.q
        inv_R = 1.0 / self.R_total
        NS = kn.mstep_cd_nstats(q)
        single = not (self.reducer.active and self.reducer.world_size > 1)
        if single and not cd.first_pending_extra and self.R > 0:
            call("pgpfa_mstep_cd_solve", ptr(self.y), ptr(est.x), ptr(est.vsm), self.R, q, N, self.T, inv_R,
                 float(cd.prior_w), ptr(cd.prior_mat), ptr(cd.theta0), ptr(cd.th_cur), ptr(cd.th_try), ptr(cd.fcur),
                 ptr(cd.step), ptr(cd.alpha), ptr(cd.slope), ptr(cd.done), ptr(cd.n_open), ptr(cd.stats), cd.issued + 1,
                 int(n_iters), float(cd.tol), ptr(self._cd_ws[1]), self._cd_ws[1].numel(), stream())
            cd.issued += n_iters
            return
        for _ in range(n_iters):
            it = cd.issued + 1
            gate = None if it == 1 else cd.n_open
            call("pgpfa_mstep_cd_stats_gated", ptr(self.y), ptr(est.x), ptr(est.vsm), ptr(cd.th_try), self.R, q, N, self.T,
                 ptr(cd.stats), ptr(self._cd_ws[1]), self._cd_ws[1].numel(), ptr(gate), stream())
            if cd.first_pending_extra:                    # the objective sum and the E-step error flags ride along
                cd.stats[NS:NS + 3, 0] = est.flags()
            self.reducer.sum_tensor(cd.stats)
            if cd.first_pending_extra:
                cd.extra = cd.stats[NS:NS + 3, 0].clone()
                cd.first_pending_extra = False
            call("pgpfa_mstep_cd_update", ptr(cd.stats), inv_R, float(cd.prior_w), ptr(cd.prior_mat), ptr(cd.theta0),
                 ptr(cd.th_cur), ptr(cd.th_try), ptr(cd.fcur), ptr(cd.step), ptr(cd.alpha), ptr(cd.slope), ptr(cd.done),
                 1 if it == 1 else 0, float(cd.tol), N, q, ptr(cd.n_open), it, stream())
            cd.issued += 1

    def mstep_cd_async(self, params, est, prior_w=0.0, tol=1e-10, one_step=False, step_size=1.0, prior_mat=None,
                       n_blind=4):
        """Per-neuron damped Newton on MStepObservationCost (+ prior term), enqueued without host reads:
        `n_blind` iterations (the usual solve takes 3-4; iterations after convergence are empty launches), the rest —
        if any neuron is still open — in CdSolve.finish().  With trial sharding the per-neuron statistics are
        all-reduced once per iteration; the first reduction also carries the E-step's objective sum and error code."""
        main = torch.cuda.current_stream()
        N, q = params.C.shape
        P = q + 1
        cd = CdSolve(self, est, q, N)
        cd.main, cd.one_step = main, one_step
        cd.prior_w, cd.prior_mat, cd.tol = prior_w, prior_mat, (0.0 if one_step else tol)
        with est.means_stream():
            cd.theta0 = params.theta
            cd.th_cur, cd.th_try = cd.theta0.clone(), cd.theta0.clone()
            cd.fcur, cd.alpha, cd.slope = empty(N), empty(N), empty(N)
            cd.step = empty(N, P)
            cd.done = torch.zeros(N, dtype=torch.int32, device="cuda")
            cd.n_open = torch.zeros(4, dtype=torch.int32, device="cuda")
            # per-neuron statistics + 3 extra rows whose first column carries the E-step flags through the first all-reduce
            cd.stats = torch.zeros(kn.mstep_cd_nstats(q) + 3, N, dtype=torch.float64, device="cuda")
            if self._cd_ws is None or self._cd_ws[0] != (q, N):
                self._cd_ws = ((q, N), _lib.workspace(_lib.lib.pgpfa_mstep_cd_workspace_bytes(q, N)))
            multi = self.reducer.active and self.reducer.world_size > 1
            cd.first_pending_extra = bool(multi)
            if not multi:
                cd.extra = est.flags()
            if one_step:
                # 'grad' online rule (funs/learning.py:884-891 with the analytic Hessian): one scaled Newton step
                self._cd_iterate_one_step(cd, step_size)
            else:
                self._cd_iterate(cd, n_blind)
        return cd

    def _cd_iterate_one_step(self, cd, step_size):
        est = cd.est
        N, q = cd.N, cd.q
        NS = kn.mstep_cd_nstats(q)
        inv_R = 1.0 / self.R_total
        call("pgpfa_mstep_cd_stats_gated", ptr(self.y), ptr(est.x), ptr(est.vsm), ptr(cd.th_try), self.R, q, N, self.T,
             ptr(cd.stats), ptr(self._cd_ws[1]), self._cd_ws[1].numel(), None, stream())
        if cd.first_pending_extra:
            cd.stats[NS:NS + 3, 0] = est.flags()
        self.reducer.sum_tensor(cd.stats)
        if cd.first_pending_extra:
            cd.extra = cd.stats[NS:NS + 3, 0].clone()
            cd.first_pending_extra = False
        cd.hess = cd.stats[:NS].clone()
        call("pgpfa_mstep_cd_update", ptr(cd.stats), inv_R, float(cd.prior_w), ptr(cd.prior_mat), ptr(cd.theta0),
             ptr(cd.th_cur), ptr(cd.th_try), ptr(cd.fcur), ptr(cd.step), ptr(cd.alpha), ptr(cd.slope), ptr(cd.done), 1, 0.0,
             N, q, ptr(cd.n_open), 1, stream())
        cd.th_cur = kn.emap("axpy", cd.theta0, cd.step, step_size)
        cd.n_open.zero_()
        cd.issued = 1

    def mstep_cd(self, params, est, prior_w=0.0, tol=1e-10, max_iter=100, one_step=False, step_size=1.0,
                 prior_mat=None):
        """Returns (C, d, cost, iterations, hess) — see mstep_cd_async."""
        cd = self.mstep_cd_async(params, est, prior_w=prior_w, tol=tol, one_step=one_step, step_size=step_size,
                                 prior_mat=prior_mat, n_blind=min(4, max_iter))
        out = cd.finish(max_iter=max_iter)
        if not est._checked:
            EStepResult.raise_on(cd.extra_host)
            est._checked = True
        return out

    def cd_cost_grad(self, theta, est):
        """(cost, grad (N,q+1)) of MStepObservationCost at theta, normalised by the global trial count."""
        q = theta.shape[1] - 1
        if self.R > 0:
            stats = kn.mstep_cd_stats(self.y, est.x, est.vsm, theta)
        else:
            stats = torch.zeros(kn.mstep_cd_nstats(q), self.N, dtype=torch.float64, device="cuda")
        stats = self.reducer.sum_tensor(stats)
        return float(_lib.to_host(stats[0].sum())) / self.R_total, (stats[1:q + 2].T / self.R_total).contiguous(), stats

    # ------------------------------------------------------------------ M-step tau
    def pautosum(self, est):
        """PautoSum over ALL trials (makePrecomp, funs/learning.py:145-173): this rank's sum — taken inside the low-rank
        posterior pass when the E-step was asked for it, else over the per-trial blocks — all-reduced over the ranks."""
        if getattr(est, "pautosum", None) is not None:
            P = est.pautosum.clone()
        elif self.R > 0:
            P = kn.pautosum(est.get_vsmGP() if hasattr(est, "get_vsmGP") else est.vsmGP, est.x)
        else:
            P = torch.zeros(est.params.q, self.T, self.T, dtype=torch.float64, device="cuda")
        return self.reducer.sum_tensor(P)

    def _tau_rounds(self, ts, n_rounds):
        call("pgpfa_mstep_tau_solve", ptr(ts.Psum), ptr(ts.tau_old), float(ts.R), ts.q, self.T, EPS_NOISE, float(ts.pw),
             self.binSize, float(ts.xtol), ts.m, ts.rounds, int(n_rounds), ptr(ts.tau), ptr(ts.det), ptr(ts.flags),
             ptr(self._tau_ws[1]), self._tau_ws[1].numel(), stream())
        ts.rounds += n_rounds

    def mstep_tau_async(self, params, Psum, numTrials=None, prior_step=None, xtol=1e-10, ncand=9, n_blind=3):
        """q independent scalar minimisations over p = log(1/tau_bins^2) (funs/learning.py:257-293, :771-830), searched on
        the device (pgpfa_mstep_tau_solve, csrc/tau_search.h): every round evaluates `ncand` candidate points per latent
        in one batched launch sequence (the T x T factorisations are latency-bound, so candidates are free) and a
        controller kernel brackets the sign change of the gradient / refines it by inverse interpolation.  `n_blind`
        rounds are enqueued without host reads (the usual search takes 3); TauSolve.finish() adds rounds if a latent is
        still open.  Every rank runs the identical search on the all-reduced PautoSum."""
        q = params.q
        m = int(ncand)
        ts = TauSolve()
        ts.trials, ts.q, ts.m, ts.xtol = self, q, m, xtol
        ts.R = float(self.R_total if numTrials is None else numTrials)
        ts.pw = 0.0 if prior_step is None else 1.0 / float(prior_step) ** 2
        key = (q, self.T, m)
        if self._tau_ws is None or self._tau_ws[0] != key:
            self._tau_ws = (key, _lib.workspace(_lib.lib.pgpfa_tau_solve_workspace_bytes(q, self.T, m)))
        ts.Psum, ts.tau_old = Psum, params.tau
        ts.tau, ts.det = empty(q), empty(6, q)
        ts.flags = torch.zeros(4, dtype=torch.int32, device="cuda")
        ts.rounds = 0
        self._tau_rounds(ts, n_blind)
        return ts

    def mstep_tau(self, params, Psum, numTrials=None, prior_step=None, xtol=1e-10, max_rounds=14, ncand=9):
        """Returns (tau_seconds (q) as a device tensor, details dict) — see mstep_tau_async."""
        ts = self.mstep_tau_async(params, Psum, numTrials=numTrials, prior_step=prior_step, xtol=xtol, ncand=ncand)
        tau = ts.finish(max_rounds=max_rounds)
        return tau, ts.details()

    # ------------------------------------------------------------------ one device-resident EM iteration
    def em_step(self, params, x0=None, tol=1e-8, cd_tol=1e-10, tau_xtol=1e-10, inference='laplace', lam0=None):
        """One full batch EM iteration (funs/engine.py:180-239: Laplace E-step, C,d M-step, timescale M-step) with the
        iteration loops driven from the device.  Host synchronisations: one wait inside the E-step driver (the number of
        trials that need the exact-Newton fallback decides what is enqueued next) and ONE packed read at the end
        (M-step flags, objective, error codes, and — for the next iteration's prior, already enqueued — positive
        definiteness and the low-rank ranks).  Returns (new DeviceParams, EStepResult, post_lik, info dict)."""
        if inference == 'laplace':
            est = self.estep_laplace(params, x0=x0, tol=tol, want_vsmGP=False, want_pautosum=True)
        else:       # dual variational E-step (funs/engine.py:202-209): same M-step on the variational posterior
            est = self.estep_variational(params, lam0=lam0)
        cd = self.mstep_cd_async(params, est, tol=cd_tol)
        Psum = self.pautosum(est)
        ts = self.mstep_tau_async(params, Psum, xtol=tau_xtol, n_blind=self._tau_blind)
        cd.join()
        q = params.q
        newp = DeviceParams(cd.th_cur[:, :q].contiguous(), cd.th_cur[:, q].contiguous(), ts.tau, self.T, self.binSize)
        pend_cd, pend_ts, pend_p = cd.pending(), ts.pending(), newp.pending()
        cd.join()
        vals = read_packed(pend_cd + pend_ts + pend_p)
        v_cd, v_ts, v_p = vals[:len(pend_cd)], vals[len(pend_cd):len(pend_cd) + len(pend_ts)], vals[len(pend_cd) + len(pend_ts):]
        EStepResult.raise_on(v_cd[2], "Laplace" if inference == 'laplace' else "variational")
        est._checked = True
        redo = v_cd[0][0] > 0 or v_ts[0][0] > 0
        C, d, cost, cd_it, _ = cd.finish(v_cd)
        tau = ts.finish(v_ts)
        self._tau_blind = max(2, min(4, int(ts.flags_host[1])))       # steady-state EM: two rounds
        if redo:                      # a blind schedule was too short (rare): the prior was built from unfinished values
            newp = DeviceParams(C, d, tau, self.T, self.binSize)
        else:
            newp.resolve(v_p)
        lik = -float(v_cd[2][0]) / self.R_total
        return newp, est, lik, {"cd_iters": cd_it, "tau_evals": int(ts.flags_host[1]), "cd_cost": cost, "tau_solve": ts}
