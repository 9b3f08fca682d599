"""Thin tensor-level wrappers over the C ABI (one Python function per pgpfa_* entry point).

Inputs/outputs are float64 CUDA tensors in the device layouts of include/pgpfa_b200.h.  No
arithmetic happens here: each function allocates outputs/workspaces with torch and makes one call.
"""
import ctypes

import torch

from . import _lib
from ._lib import call, empty, handle, ptr, stream, workspace

QMAX = 12


def emap(op, x, y=None, a=0.0):
    """Element-wise helper on the device: op 'exp', 'log', or 'axpy' (x + a*y)."""
    code = {"exp": 0, "log": 1, "axpy": 2}[op]
    out = torch.empty_like(x)
    call("pgpfa_map", code, x.numel(), ptr(x), ptr(y), float(a), ptr(out), stream())
    return out


def bin_spikes(times, ptr_, t0, dur, R, N, T):
    """Spike times (float64, CSR offsets int64 over (trial, neuron)) -> counts (R,N,T), numpy.histogram rules."""
    Y = empty(R, N, T)
    call("pgpfa_bin_spikes", ptr(times) if times.numel() else None, ptr(ptr_), ptr(t0), float(dur), R, N, T, ptr(Y), stream())
    return Y


def sample_dataset(C, d, tau, R, T, binSize, seed, epsNoise=0.001):
    """Draw X (R,q,T) ~ GP(0,K(tau)) and Y (R,N,T) ~ Poisson(exp(CX+d)) on the device (funs/util.py:733-752)."""
    q = tau.numel()
    N = C.shape[0]
    K = make_K(tau, T, binSize, epsNoise)
    L, D, _, info = potrf_dense(K, want_zt=False)
    Ld = tiles_to_dense(L, T)
    z = empty(R, q, T)
    call("pgpfa_sample_normal", ptr(z), z.numel(), int(seed) & 0xFFFFFFFFFFFFFFFF, stream())
    X = prior_apply(Ld, z)                                          # x_k = chol(K_k) z_k
    Y = empty(R, N, T)
    call("pgpfa_sample_poisson", ptr(X), ptr(C), ptr(d), R, q, N, T, (int(seed) * 0x9E3779B97F4A7C15 + 1) & 0xFFFFFFFFFFFFFFFF,
         ptr(Y), stream())
    return X, Y


def make_K(tau, T, binSize, epsNoise=0.001):
    """funs/util.py:599-614 -> K (q,T,T)."""
    q = tau.numel()
    K = empty(q, T, T)
    call("pgpfa_make_K", ptr(tau), q, T, float(binSize), float(epsNoise), ptr(K), stream())
    return K


def make_K_big(K):
    """funs/util.py:616-617 block-diagonal assembly."""
    q, T, _ = K.shape
    Kb = empty(q * T, q * T)
    call("pgpfa_make_K_big", ptr(K), q, T, ptr(Kb), stream())
    return Kb


def make_K_gamma(p, T, epsNoise=0.001, want_dK=True):
    q = p.numel()
    K = empty(q, T, T)
    dK = empty(q, T, T) if want_dK else None
    call("pgpfa_make_K_gamma", ptr(p), q, T, float(epsNoise), ptr(K), ptr(dK), stream())
    return K, dK


def spd_inverse(A):
    """Batched SPD inverse and log-determinant: A (b,n,n) -> (Ainv, logdet (b), info (b))."""
    b, n, _ = A.shape
    Ainv = empty(b, n, n)
    logdet = empty(b)
    info = empty(b, dtype=torch.int32)
    nbytes = _lib.lib.pgpfa_spd_inverse_workspace_bytes(b, n)
    ws = workspace(nbytes)
    call("pgpfa_spd_inverse_batched", ptr(A), b, n, ptr(Ainv), ptr(logdet), ptr(info), ptr(ws), nbytes, stream())
    return Ainv, logdet, info


def tile_buffers(batch, n, want_zt=True):
    tb = _lib.lib.pgpfa_tiles_bytes(n) // 8
    db = _lib.lib.pgpfa_dinv_bytes(n) // 8
    L = empty(batch, tb)
    D = empty(batch, db)
    ZT = empty(batch, tb) if want_zt else None
    return L, D, ZT


def potrf_dense(A, want_zt=True):
    b, n, _ = A.shape
    L, D, ZT = tile_buffers(b, n, want_zt)
    info = empty(b, dtype=torch.int32)
    call("pgpfa_potrf_dense", ptr(A), b, n, ptr(L), ptr(D), ptr(ZT), ptr(info), stream())
    return L, D, ZT, info


def potrf_posterior(Kinv, W, diag_scale=1.0, want_zt=True, bufs=None):
    """Factor H_r = blkdiag(Kinv) + scatter(W_r) for all r without materialising it."""
    q, T, _ = Kinv.shape
    b = W.shape[0]
    L, D, ZT = bufs if bufs is not None else tile_buffers(b, q * T, want_zt)
    info = empty(b, dtype=torch.int32)
    call("pgpfa_potrf_posterior", ptr(Kinv), ptr(W), float(diag_scale), b, q, T, ptr(L), ptr(D), ptr(ZT), ptr(info),
         stream())
    return L, D, ZT, info


def potrs(L, D, rhs, scale=1.0):
    b, n = rhs.shape[0], rhs[0].numel()
    out = torch.empty_like(rhs)
    call("pgpfa_potrs", ptr(L), ptr(D), ptr(rhs), float(scale), b, n, ptr(out), stream())
    return out


def trtri(L, D, ZT, n):
    call("pgpfa_trtri", ptr(L), ptr(D), ptr(ZT), L.shape[0], n, stream())
    return ZT


def potri_dense(ZT, n):
    b = ZT.shape[0]
    out = empty(b, n, n)
    ws = workspace(((n + 63) // 64) ** 2 * 8 + 1024)
    call("pgpfa_potri_dense", ptr(ZT), b, n, ptr(out), ptr(ws), ws.numel(), stream())
    return out


def cov_slices(ZT, q, T, want_vsm=True, want_vsmGP=True):
    b = ZT.shape[0]
    vsm = empty(b, T, q, q) if want_vsm else None
    vsmGP = empty(b, q, T, T) if want_vsmGP else None
    ws = workspace(((q * T + 63) // 64) ** 2 * 8 + 1024)
    call("pgpfa_cov_slices", ptr(ZT), b, q, T, ptr(vsm), ptr(vsmGP), ptr(ws), ws.numel(), stream())
    return vsm, vsmGP


def logdet(L, n):
    out = empty(L.shape[0])
    call("pgpfa_logdet", ptr(L), L.shape[0], n, ptr(out), stream())
    return out


def tiles_to_dense(tiles, n, upper=False):
    out = empty(tiles.shape[0], n, n)
    call("pgpfa_tiles_to_dense", ptr(tiles), tiles.shape[0], n, int(upper), ptr(out), stream())
    return out


def prior_apply(Kmat, v):
    R, q, T = v.shape
    out = torch.empty_like(v)
    call("pgpfa_prior_apply", ptr(Kmat), ptr(v), R, q, T, ptr(out), stream())
    return out


def laplace_eval(x, y, C, d, Kinv):
    """(f (R), g (R,q,T), W (R,q*q,T)) of funs/inference.py:12-65 at x (R,q,T)."""
    R, q, T = x.shape
    N = y.shape[1]
    f, g, W, Kx = empty(R), empty(R, q, T), empty(R, q * q, T), empty(R, q, T)
    call("pgpfa_laplace_eval", ptr(x), ptr(y), ptr(C), ptr(d), ptr(Kinv), R, q, N, T, ptr(f), ptr(g), ptr(W), ptr(Kx),
         stream())
    return f, g, W


def hessian_dense(Kinv, W, diag_scale=1.0):
    q, T, _ = Kinv.shape
    R = W.shape[0]
    H = empty(R, q * T, q * T)
    call("pgpfa_hessian_dense", ptr(Kinv), ptr(W), float(diag_scale), R, q, T, ptr(H), stream())
    return H


class LaplaceResult:
    __slots__ = ("x", "f", "vsm", "vsmGP", "cov", "niter", "info", "stats", "rc", "pautosum")


def prior_lowrank_async(K, eps=0.001, delta=1e-14, out=None):
    """Pivoted Cholesky K_k - eps I = F_k F_k^T (pgpfa_prior_lowrank), enqueued only: (F, Ft, ranks as a DEVICE tensor).
    `out` = preallocated (F, Ft, rank) when the call is made under another stream than the one that owns the memory."""
    q, T, _ = K.shape
    if out is None:
        out = (empty(q, T, T), empty(q, T, T), empty(q, dtype=torch.int32))
    F, Ft, rank = out
    call("pgpfa_prior_lowrank", ptr(K), q, T, float(eps), float(delta), ptr(F), ptr(Ft), ptr(rank), stream())
    return F, Ft, rank


def prior_lowrank(K, eps=0.001, delta=1e-14):
    """Same with the ranks read back: (F, Ft, ranks as a host list)."""
    F, Ft, rank = prior_lowrank_async(K, eps, delta)
    return F, Ft, [int(v) for v in _lib.to_host(rank)]


def laplace_solve(y, C, d, Kinv, x0=None, tol=1e-8, max_newton=50, want_vsm=True, want_vsmGP=True, want_cov=False,
                  max_ws_bytes=None, ws=None, inexact_newton=True, lowrank=None, want_pautosum=False):
    """Batched Newton E-step (pgpfa_laplace_solve). y (R,N,T); returns LaplaceResult with device tensors.
    lowrank = (F, Ft, ranks, eps) from prior_lowrank: posterior pass through the low-rank prior factor
    (pgpfa_laplace_solve_lowrank; not with want_cov).  want_pautosum (low-rank pass only): res.pautosum (q,T,T) = the
    trial-sum of post_vsmGP + m m^T computed inside the pass; together with want_vsmGP=False the per-trial blocks are
    never written."""
    R, N, T = y.shape
    q = C.shape[1]
    x = torch.zeros(R, q, T, dtype=torch.float64, device="cuda") if x0 is None else x0.clone()
    res = LaplaceResult()
    res.x = x
    res.f = empty(R)
    res.vsm = empty(R, T, q, q) if want_vsm else None
    res.vsmGP = empty(R, q, T, T) if want_vsmGP else None
    res.pautosum = None
    res.cov = empty(R, q * T, q * T) if want_cov else None
    res.niter = empty(R, dtype=torch.int32)
    res.info = empty(R, dtype=torch.int32)
    full = _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, R)
    if ws is None:
        nbytes = full
        if max_ws_bytes is None:
            free, _ = torch.cuda.mem_get_info()
            max_ws_bytes = int(free * 0.85)
        if nbytes > max_ws_bytes:
            nbytes = max(max_ws_bytes, _lib.lib.pgpfa_laplace_workspace_bytes(R, q, T, 1))
        ws = workspace(nbytes)
    stats = (ctypes.c_int * 8)()
    if lowrank is not None and not want_cov:
        F, Ft, ranks, eps = lowrank
        rank_host = (ctypes.c_int * q)(*[int(v) for v in ranks])
        if want_pautosum:
            res.pautosum = empty(q, T, T)
        res.rc = call("pgpfa_laplace_solve_lowrank", handle(), ptr(y), ptr(C), ptr(d), ptr(Kinv), ptr(F), ptr(Ft),
                      ctypes.cast(rank_host, ctypes.c_void_p), float(eps), ptr(x), R, q, N, T, float(tol), int(max_newton),
                      int(bool(inexact_newton)), ptr(res.f), ptr(res.vsm), ptr(res.vsmGP), ptr(res.pautosum), ptr(res.niter),
                      ptr(res.info),
                      ptr(ws), ws.numel(), ctypes.cast(stats, ctypes.c_void_p), stream(), allow=(_lib.ERR_NOT_CONVERGED,))
    else:
        if want_pautosum and res.vsmGP is None:          # dense tiled path: the sum is taken over the per-trial blocks
            res.vsmGP = empty(R, q, T, T)
        res.rc = call("pgpfa_laplace_solve", handle(), ptr(y), ptr(C), ptr(d), ptr(Kinv), ptr(x), R, q, N, T, float(tol),
                      int(max_newton), int(bool(inexact_newton)), ptr(res.f), ptr(res.vsm), ptr(res.vsmGP), ptr(res.cov),
                      ptr(res.niter), ptr(res.info), ptr(ws), ws.numel(), ctypes.cast(stats, ctypes.c_void_p), stream(),
                      allow=(_lib.ERR_NOT_CONVERGED,))
        if want_pautosum:
            res.pautosum = pautosum(res.vsmGP, res.x)
    res.stats = {"factorizations": stats[0], "max_newton_iters": stats[1], "not_converged": stats[2],
                 "chunk": stats[3], "pcg_newton_iters": stats[4], "fallback_trials": stats[5], "lowrank_r": stats[6],
                 "fresh_chord_sweeps": stats[7] % 1000, "pcg_iters": stats[7] // 1000}
    return res


def dualvi_eval(lam, y, C, d, K, Kinv, want_grad=True, want_cov=False):
    """dualProblem / dualProblem_grad / VIPostMean / VIPostCov slices at lambda (R,N,T) — funs/inference.py:188-219."""
    R, N, T = y.shape
    q = C.shape[1]
    D, mean, vsm = empty(R), empty(R, q, T), empty(R, T, q, q)
    grad = empty(R, N, T) if want_grad else None
    cov = empty(R, q * T, q * T) if want_cov else None
    info = empty(R, dtype=torch.int32)
    nbytes = _lib.lib.pgpfa_dualvi_workspace_bytes(R, q, T, R)
    ws = workspace(nbytes)
    call("pgpfa_dualvi_eval", handle(), ptr(lam), ptr(y), ptr(C), ptr(d), ptr(K), ptr(Kinv), R, q, N, T, ptr(D),
         ptr(grad), ptr(mean), ptr(vsm), ptr(cov), ptr(info), ptr(ws), nbytes, stream())
    return D, grad, mean, vsm, cov


def rate_blocks(lam, C):
    """W (R, q*q, T) with W[r,k*q+l,t] = sum_n C[n,k] C[n,l] lam[r,n,t]."""
    R, N, T = lam.shape
    q = C.shape[1]
    W = empty(R, q * q, T)
    scratch = empty(R * q * T + 2 * R)
    call("pgpfa_rate_blocks", ptr(lam), ptr(C), R, q, N, T, ptr(W), ptr(scratch), stream())
    return W


class DualVIResult:
    __slots__ = ("x", "s", "lam", "mean", "D", "f", "vsm", "vsmGP", "cov", "niter", "info", "stats", "rc")


def dualvi_solve(y, C, d, K, Kinv, lam0=None, tol=1e-10, max_iter=200, want_vsmGP=True, want_cov=False, ws=None):
    """Fixed point of the dual problem for all trials (pgpfa_dualvi_solve)."""
    R, N, T = y.shape
    q = C.shape[1]
    res = DualVIResult()
    res.x = torch.zeros(R, q, T, dtype=torch.float64, device="cuda")
    res.s = torch.zeros(R, N, T, dtype=torch.float64, device="cuda")
    nbytes = _lib.lib.pgpfa_dualvi_workspace_bytes(R, q, T, R)
    if ws is None:
        free, _ = torch.cuda.mem_get_info()
        budget = int(free * 0.8)
        if nbytes > budget:
            nbytes = max(budget, _lib.lib.pgpfa_dualvi_workspace_bytes(R, q, T, 1))
        ws = workspace(nbytes)
    if lam0 is not None:
        call("pgpfa_dualvi_init_from_lambda", ptr(lam0), ptr(y), ptr(C), ptr(d), ptr(K), R, q, N, T, ptr(res.x),
             ptr(res.s), ptr(ws), ws.numel(), stream())
    res.lam, res.mean, res.D, res.f = empty(R, N, T), empty(R, q, T), empty(R), empty(R)
    res.vsm = empty(R, T, q, q)
    res.vsmGP = empty(R, q, T, T) if want_vsmGP else None
    res.cov = empty(R, q * T, q * T) if want_cov else None
    res.niter, res.info = empty(R, dtype=torch.int32), empty(R, dtype=torch.int32)
    stats = (ctypes.c_int * 4)()
    res.rc = call("pgpfa_dualvi_solve", handle(), ptr(y), ptr(C), ptr(d), ptr(K), ptr(Kinv), ptr(res.x), ptr(res.s),
                  R, q, N, T, float(tol), int(max_iter), ptr(res.lam), ptr(res.mean), ptr(res.D), ptr(res.f),
                  ptr(res.vsm), ptr(res.vsmGP), ptr(res.cov), ptr(res.niter), ptr(res.info), ptr(ws), ws.numel(),
                  ctypes.cast(stats, ctypes.c_void_p), stream(), allow=(_lib.ERR_NOT_CONVERGED,))
    res.stats = {"factorizations": stats[0], "sweeps": stats[1], "not_converged": stats[2], "chunk": stats[3]}
    return res


def loo_predict(y, C, d, Kinv, ymap, excl, tol=1e-8, max_newton=60):
    """Leave-one-neuron-out prediction for problems p = (trial ymap[p], neuron excl[p]) (int32 device tensors).
    Returns (ypred (P,T), err (P), modes (P,q,T), stats)."""
    R, N, T = y.shape
    q = C.shape[1]
    Pn = ymap.numel()
    x = torch.zeros(Pn, q, T, dtype=torch.float64, device="cuda")
    ypred, err = empty(Pn, T), empty(Pn)
    niter, info = empty(Pn, dtype=torch.int32), empty(Pn, dtype=torch.int32)
    full = _lib.lib.pgpfa_laplace_workspace_bytes(Pn, q, T, Pn)
    free, _ = torch.cuda.mem_get_info()
    nbytes = full if full <= int(free * 0.8) else max(int(free * 0.8), _lib.lib.pgpfa_laplace_workspace_bytes(Pn, q, T, 1))
    ws = workspace(nbytes)
    stats = (ctypes.c_int * 8)()
    rc = call("pgpfa_loo_predict", handle(), ptr(y), ptr(C), ptr(d), ptr(Kinv), ptr(ymap), ptr(excl), ptr(x), Pn, q, N, T,
              float(tol), int(max_newton), ptr(ypred), ptr(err), ptr(niter), ptr(info), ptr(ws), ws.numel(),
              ctypes.cast(stats, ctypes.c_void_p), stream(), allow=(_lib.ERR_NOT_CONVERGED,))
    return ypred, err, x, {"rc": rc, "factorizations": stats[0], "max_newton_iters": stats[1], "not_converged": stats[2]}


def pautosum(vsmGP, post_mean, out=None, accumulate=False):
    R, q, T, _ = vsmGP.shape
    P = out if out is not None else empty(q, T, T)
    call("pgpfa_pautosum", ptr(vsmGP), ptr(post_mean), R, q, T, int(accumulate), ptr(P), stream())
    return P


def mstep_cd_nstats(q):
    return _lib.lib.pgpfa_mstep_cd_nstats(q)


def mstep_cd_stats(y, post_mean, vsm, theta, ws=None):
    """Un-normalised per-neuron (cost, grad, Hessian) sums over the local trials: (NS, N)."""
    R, N, T = y.shape
    q = post_mean.shape[1]
    stats = empty(mstep_cd_nstats(q), N)
    nbytes = _lib.lib.pgpfa_mstep_cd_workspace_bytes(q, N)
    ws = workspace(nbytes) if ws is None else ws
    call("pgpfa_mstep_cd_stats", ptr(y), ptr(post_mean), ptr(vsm), ptr(theta), R, q, N, T, ptr(stats), ptr(ws),
         ws.numel(), stream())
    return stats


def tau_eval(p, Psum, numTrials, T, epsNoise=0.001, prior_w=0.0, tau_old=None, binSize=10.0, ws=None):
    q = p.numel()
    cost, grad = empty(q), empty(q)
    nbytes = _lib.lib.pgpfa_tau_eval_workspace_bytes(q, T)
    ws = workspace(nbytes) if ws is None else ws
    call("pgpfa_tau_eval", ptr(p), ptr(Psum), float(numTrials), q, T, float(epsNoise), float(prior_w), ptr(tau_old),
         float(binSize), ptr(cost), ptr(grad), ptr(ws), ws.numel(), stream())
    return cost, grad
