"""CPU oracle for the Poisson-GPFA EM hot path — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline /
``--impl reference`` legs may import this file.  It is a numpy/scipy restatement of
the reference algorithm (mackelab/poisson-gpfa, ``funs/*.py``); every function cites
the reference lines it follows.  It is pinned against the real reference by
``oracle/validate_oracle.py`` (run in the build container, where /root/reference
exists) and against the golden vectors in ``tests/golden`` (generated from the real
reference by ``oracle/make_golden.py``).

Two flavours live side by side:

* ``dense_*``   — the reference's own formulation: materialises C_big (qT x NT) and
  K_big, calls the same scipy optimisers with the same options.  This is the
  "port" that is timed as the CPU baseline, because it does the same work the
  reference does on a CPU.
* ``*_struct``  — the same mathematics using the per-bin identities of SURVEY.md
  §8(a) (no C_big), with exact Newton / tight optimiser tolerances.  Used as the
  checker at sizes where the dense formulation would take minutes.

Layouts (funs/util.py:594-597, funs/inference.py:97): xbar[k*T+t] = x[k,t],
ybar[n*T+t] = y[n,t]; vecCd[j*N+n] = C[n,j], vecCd[q*N+n] = d[n]
(funs/util.py:560-592).
"""
import copy

import numpy as np
import scipy.optimize as sopt

EPS_NOISE = 0.001  # funs/util.py:599


# ----------------------------------------------------------------------------
# GP prior and big-matrix builders
# ----------------------------------------------------------------------------
def make_K(tau, T, binSize, epsNoise=EPS_NOISE):
    """K[k,i,j] = (1-eps) exp(-0.5 ((i-j) binSize)^2 / (tau_k 1000)^2) + eps delta_ij.
    funs/util.py:599-614 (vectorised; same operation order inside the exp)."""
    tau = np.asarray(tau, dtype=np.float64).ravel()
    t_ms = np.arange(T) * binSize
    dif = t_ms[:, None] - t_ms[None, :]
    K = np.empty((tau.size, T, T))
    for k in range(tau.size):
        K[k] = (1.0 - epsNoise) * np.exp(-0.5 * (dif ** 2 / (tau[k] * 1000) ** 2))
        K[k] += epsNoise * np.eye(T)
    return K


def make_K_big(params, trialDur, binSize, epsNoise=EPS_NOISE):
    """(K_big, K) as funs/util.py:599-619, including the in-place flatten of params['tau'] (:602)."""
    params['tau'] = np.ndarray.flatten(np.asarray(params['tau'], dtype=np.float64))
    T = int(trialDur / binSize)
    K = make_K(params['tau'], T, binSize, epsNoise)
    q = K.shape[0]
    K_big = np.zeros((q * T, q * T))
    for k in range(q):
        K_big[k * T:(k + 1) * T, k * T:(k + 1) * T] = K[k]
    return K_big, K


def make_Cd_big(params, T):
    """funs/util.py:594-597."""
    C_big = np.kron(params['C'], np.eye(T)).T
    d_big = np.kron(np.ravel(params['d']), np.ones(T)).T
    return C_big, d_big


def Cd_to_vec(C, d):
    """funs/util.py:560-574: latent-major stacking, d last."""
    return np.concatenate([np.asarray(C).T.ravel(), np.ravel(d)])


def vec_to_Cd(vecCd, xdim, ydim):
    """funs/util.py:576-592."""
    m = np.reshape(vecCd, (xdim + 1, ydim)).T
    return m[:, :xdim], m[:, xdim]


# ----------------------------------------------------------------------------
# Laplace objective: dense (reference formulation) and structured
# ----------------------------------------------------------------------------
def dense_nlp(xbar, ybar, C_big, d_big, K_bigInv):
    """funs/inference.py:12-32."""
    a = C_big.T @ xbar + d_big
    return np.exp(a).sum() - ybar @ a + 0.5 * xbar @ (K_bigInv @ xbar)


def dense_nlp_grad(xbar, ybar, C_big, d_big, K_bigInv):
    """funs/inference.py:34-48."""
    lam = np.exp(C_big.T @ xbar + d_big)
    return lam @ C_big.T - ybar @ C_big.T + xbar @ K_bigInv


def dense_nlp_hess(xbar, ybar, C_big, d_big, K_bigInv):
    """funs/inference.py:50-65 (diag(lam) applied as a row scaling)."""
    lam = np.exp(C_big.T @ xbar + d_big)
    return C_big @ (lam[:, None] * C_big.T) + K_bigInv


def nlp_struct(x, y, C, d, Kinv):
    """Same value as dense_nlp using per-bin identities. x (q,T), y (N,T), Kinv (q,T,T)."""
    h = C @ x + d[:, None]
    prior = 0.5 * np.einsum('kt,kts,ks->', x, Kinv, x)
    return np.exp(h).sum() - (y * h).sum() + prior


def nlp_grad_struct(x, y, C, d, Kinv):
    lam = np.exp(C @ x + d[:, None])
    return C.T @ (lam - y) + np.einsum('kts,ks->kt', Kinv, x)


def nlp_W_struct(x, C, d):
    """W[t,k,l] = sum_n C[n,k] C[n,l] lam[n,t]  (the per-bin blocks of C_big diag(lam) C_big^T)."""
    lam = np.exp(C @ x + d[:, None])
    return np.einsum('nk,nl,nt->tkl', C, C, lam)


def assemble_H(Kinv, W, diag_jitter=0.0):
    """H = blkdiag(Kinv_k) + scatter(W) in latent-major ordering; optional relative diagonal jitter
    (funs/inference.py:190 adds 1e-6*diag(diag(P)) in the variational path)."""
    q, T, _ = Kinv.shape
    H = np.zeros((q * T, q * T))
    for k in range(q):
        H[k * T:(k + 1) * T, k * T:(k + 1) * T] = Kinv[k]
    idx = np.arange(T)
    for k in range(q):
        for l in range(q):
            H[k * T + idx, l * T + idx] += W[:, k, l]
    if diag_jitter:
        H[np.diag_indices_from(H)] *= (1.0 + diag_jitter)
    return H


def slice_cov(cov, q, T):
    """post_vsmGP (T,T,q) and post_vsm (T,q,q) slices, funs/inference.py:164-172."""
    vsmGP = np.zeros((T, T, q))
    for k in range(q):
        vsmGP[:, :, k] = cov[k * T:(k + 1) * T, k * T:(k + 1) * T]
    vsm = np.zeros((T, q, q))
    for t in range(T):
        vsm[t] = cov[t::T, t::T]
    return vsmGP, vsm


def pivoted_cholesky(S, delta=1e-14):
    """S ~= F F^T by greedy pivoted Cholesky, stopped when the largest residual diagonal entry is <= delta.
    Restates csrc/lowrank.cu:pivchol_kernel (test infrastructure for the low-rank posterior pass)."""
    T = S.shape[0]
    dres = np.diag(S).astype(np.float64).copy()
    F = np.zeros((T, T))
    r = 0
    while r < T:
        j = int(np.argmax(dres))
        if not dres[j] > delta:
            break
        col = (S[:, j] - F[:, :r] @ F[j, :r]) / np.sqrt(dres[j])
        F[:, r] = col
        dres = np.maximum(dres - col ** 2, 0.0)
        dres[j] = -1.0
        r += 1
    return F[:, :r]


def lowrank_posterior_slices(x, C, d, K, epsNoise=EPS_NOISE, delta=1e-14):
    """post_vsmGP (T,T,q), post_vsm (T,q,q) of funs/inference.py:164-172 WITHOUT the qT x qT inverse: with
    K_k - eps I = F_k F_k^T, P_t = (I + eps W_t)^-1 and Dt_t = W_t P_t,  Sigma = eps P + Y Y^T,
    Y = P F L^-T,  L L^T = I + F^T Dt F  (Woodbury; the identity behind csrc/lowrank.cu)."""
    q, T, _ = K.shape
    W = nlp_W_struct(x, C, d)                                   # (T,q,q)
    Fs = [pivoted_cholesky(K[k] - epsNoise * np.eye(T), delta) for k in range(q)]
    off = np.concatenate([[0], np.cumsum([f.shape[1] for f in Fs])]).astype(int)
    r = int(off[-1])
    P = np.stack([np.linalg.inv(np.eye(q) + epsNoise * W[t]) for t in range(T)])
    Dt = np.einsum('tkl,tlm->tkm', W, P)
    G = np.eye(r)
    for k in range(q):
        for l in range(q):
            G[off[k]:off[k + 1], off[l]:off[l + 1]] += Fs[k].T @ (Dt[:, k, l][:, None] * Fs[l])
    Z = np.linalg.inv(np.linalg.cholesky(G))
    Yh = [Fs[l] @ Z[:, off[l]:off[l + 1]].T for l in range(q)]  # (T,r) per latent
    Y = [sum(P[:, k, l][:, None] * Yh[l] for l in range(q)) for k in range(q)]
    vsmGP = np.zeros((T, T, q))
    for k in range(q):
        vsmGP[:, :, k] = Y[k] @ Y[k].T + np.diag(epsNoise * P[:, k, k])
    vsm = epsNoise * P + np.einsum('ktc,ltc->tkl', np.stack(Y), np.stack(Y))
    return vsmGP, vsm, r


def _trial_list(experiment):
    return [np.asarray(tr['Y'], dtype=np.float64) for tr in experiment.data]


def dense_laplace(experiment, params, prevOptimRes=None, xtol=None):
    """The reference E-step as written: funs/inference.py:67-185 (Newton-CG, np.linalg.inv).
    ``xtol`` None keeps scipy's default 1e-5 (what the reference runs with)."""
    ys = _trial_list(experiment)
    N, T = ys[0].shape
    q = params['C'].shape[1]
    C_big, d_big = make_Cd_big(params, T)
    K_big, _ = make_K_big(params, experiment.trialDur, experiment.binSize)
    K_bigInv = np.linalg.inv(K_big)
    res = {'post_mean': [], 'post_cov': [], 'post_vsm': [], 'post_vsmGP': []}
    optim, tot = [], 0.0
    opts = {'disp': False, 'maxiter': 10000}
    if xtol is not None:
        opts['xtol'] = xtol
    for r, y in enumerate(ys):
        ybar = y.reshape(N * T)
        x0 = np.zeros(q * T) if prevOptimRes is None else prevOptimRes[r]
        out = sopt.minimize(dense_nlp, x0, args=(ybar, C_big, d_big, K_bigInv), method='Newton-CG',
                            jac=dense_nlp_grad, hess=dense_nlp_hess, options=opts)
        optim.append(out.x)
        tot += out.fun
        cov = np.linalg.inv(dense_nlp_hess(out.x, ybar, C_big, d_big, K_bigInv))
        vsmGP, vsm = slice_cov(cov, q, T)
        res['post_mean'].append(out.x.reshape(q, T))
        res['post_cov'].append(cov)
        res['post_vsm'].append(vsm)
        res['post_vsmGP'].append(vsmGP)
    return res, -tot / len(ys), optim


def newton_mode_struct(y, C, d, Kinv, x0=None, tol=1e-13, maxit=100):
    """Exact damped Newton on the strictly convex Laplace objective (the fixed point the
    reference's Newton-CG approaches, funs/inference.py:119-126). Returns (x (q,T), f, iters)."""
    q, T = Kinv.shape[0], Kinv.shape[1]
    x = np.zeros((q, T)) if x0 is None else np.array(x0, dtype=np.float64).reshape(q, T)
    f = nlp_struct(x, y, C, d, Kinv)
    it = 0
    for it in range(1, maxit + 1):
        g = nlp_grad_struct(x, y, C, d, Kinv)
        H = assemble_H(Kinv, nlp_W_struct(x, C, d))
        step = -np.linalg.solve(H, g.ravel()).reshape(q, T)
        slope = float((g * step).sum())
        a = 1.0
        while True:
            xn = x + a * step
            with np.errstate(over='ignore'):
                fn = nlp_struct(xn, y, C, d, Kinv)
            # near the optimum the decrease is below the rounding noise of f: accept the full step
            if abs(slope) <= 1e-9 * (1.0 + abs(f)):
                break
            if np.isfinite(fn) and fn <= f + 1e-4 * a * slope + 1e-13 * (1.0 + abs(f)):
                break
            a *= 0.5
            if a < 1e-10:
                break
        x, f = xn, fn
        if a == 1.0 and np.abs(step).max() <= tol * (1.0 + np.abs(x).max()):
            break
    return x, f, it


def laplace_struct(ys, params, T, binSize, x0s=None, tol=1e-13, want_cov=True):
    """Structured, tightly converged Laplace E-step on a list of (N,T) arrays."""
    C = np.asarray(params['C'], dtype=np.float64)
    d = np.ravel(np.asarray(params['d'], dtype=np.float64))
    q = C.shape[1]
    K = make_K(params['tau'], T, binSize)
    Kinv = np.stack([np.linalg.inv(K[k]) for k in range(q)])
    res = {'post_mean': [], 'post_cov': [], 'post_vsm': [], 'post_vsmGP': []}
    tot, optim, iters = 0.0, [], []
    for r, y in enumerate(ys):
        x, f, it = newton_mode_struct(y, C, d, Kinv, None if x0s is None else x0s[r], tol)
        cov = np.linalg.inv(assemble_H(Kinv, nlp_W_struct(x, C, d)))
        vsmGP, vsm = slice_cov(cov, q, T)
        res['post_mean'].append(x)
        res['post_cov'].append(cov if want_cov else None)
        res['post_vsm'].append(vsm)
        res['post_vsmGP'].append(vsmGP)
        tot += f
        optim.append(x.ravel().copy())
        iters.append(it)
    return res, -tot / len(ys), optim, iters


def leave_one_out_struct(ys, params, T, binSize, tol=1e-13):
    """funs/engine.py:599-644 with exact Newton: y_pred[r][n] = exp(c_n x_mode(-n) + d_n), summed squared error."""
    C = np.asarray(params['C'], dtype=np.float64)
    d = np.ravel(np.asarray(params['d'], dtype=np.float64))
    N, q = C.shape
    K = make_K(params['tau'], T, binSize)
    Kinv = np.stack([np.linalg.inv(K[k]) for k in range(q)])
    pred = np.zeros((len(ys), N, T))
    err = 0.0
    for r, y in enumerate(ys):
        for n in range(N):
            keep = np.arange(N) != n
            x, _, _ = newton_mode_struct(y[keep], C[keep], d[keep], Kinv, None, tol)
            pred[r, n] = np.exp(C[n] @ x + d[n])
            err += ((y[n] - pred[r, n]) ** 2).sum()
    return pred, err


# ----------------------------------------------------------------------------
# Dual variational inference
# ----------------------------------------------------------------------------
def dense_vi_post_cov(K_bigInv, C_big, lamb):
    """funs/inference.py:188-191 (row scaling instead of np.diag product)."""
    P = K_bigInv + (C_big * lamb[None, :]) @ C_big.T
    return np.linalg.inv(P + 1e-6 * np.diag(np.diag(P))), P


def dense_vi_post_mean(K_big, C_big, ybar, lamb):
    """funs/inference.py:193-194."""
    return -(K_big @ C_big) @ (lamb - ybar)


def dense_dual(lamb, ybar, C_big, K_big, K_bigInv, d_big):
    """funs/inference.py:196-213."""
    cov, _ = dense_vi_post_cov(K_bigInv, C_big, lamb)
    r = lamb - ybar
    v = C_big @ r
    A = 0.5 * v @ (K_big @ v)
    B = -d_big @ r
    _, ld = np.linalg.slogdet(cov)
    D = lamb @ (np.log(lamb) - 1.0)
    return A + B + 0.5 * ld + D


def dense_dual_grad(lamb, ybar, C_big, K_big, K_bigInv, d_big):
    """funs/inference.py:215-219; diag(C_big^T S C_big) via an einsum (same numbers, no NT x NT)."""
    cov, _ = dense_vi_post_cov(K_bigInv, C_big, lamb)
    r = lamb - ybar
    quad = np.einsum('ij,ij->j', C_big, cov @ C_big)
    return C_big.T @ (K_big @ (C_big @ r)) - d_big + np.log(lamb) - 0.5 * quad


def dense_dual_rho(rho, *a):
    """funs/inference.py:222-244."""
    return dense_dual(np.exp(rho), *a)


def dense_dual_rho_grad(rho, *a):
    """funs/inference.py:246-256."""
    return dense_dual_grad(np.exp(rho), *a) * np.exp(rho)


def dual_struct(lam, y, C, d, K, Kinv, want_grad=True):
    """Dual objective and gradient with per-bin identities. lam, y (N,T). Returns (D, grad (N,T), cov)."""
    q, T = K.shape[0], K.shape[1]
    r = lam - y
    v = C.T @ r                                   # (q,T)
    Kv = np.einsum('kts,ks->kt', K, v)
    W = np.einsum('nk,nl,nt->tkl', C, C, lam)
    P = assemble_H(Kinv, W, diag_jitter=1e-6)
    cov = np.linalg.inv(P)
    _, ld = np.linalg.slogdet(cov)
    D = 0.5 * (v * Kv).sum() - (d[:, None] * r).sum() + 0.5 * ld + (lam * (np.log(lam) - 1.0)).sum()
    if not want_grad:
        return D, None, cov
    vsm = np.stack([cov[t::T, t::T] for t in range(T)])
    quad = np.einsum('nk,tkl,nl->nt', C, vsm, C)
    g = C @ Kv - d[:, None] + np.log(lam) - 0.5 * quad
    return D, g, cov


def dual_fixed_point_struct(y, C, d, K, Kinv, tol=1e-12, maxit=1000):
    """The unique stationary point of the dual (funs/inference.py:196-219): log lam = C m + d + s,
    m = -K C_big (lam - y), s[n,t] = 0.5 c_n^T Sigma_tt c_n with Sigma = VIPostCov(lam).  Solved by the coupled
    iteration (Newton step on m with the jittered precision, refresh of s).  Returns a dict."""
    q, T = K.shape[0], K.shape[1]
    N = C.shape[0]
    m = np.zeros((q, T))
    s = np.zeros((N, T))
    for it in range(maxit):
        lam = np.exp(C @ m + d[:, None] + s)
        g = C.T @ (lam - y) + np.einsum('kts,ks->kt', Kinv, m)
        P = assemble_H(Kinv, np.einsum('nk,nl,nt->tkl', C, C, lam), diag_jitter=1e-6)
        cov = np.linalg.inv(P)
        step = -(cov @ g.ravel()).reshape(q, T)
        m = m + step
        vsm = np.stack([cov[t::T, t::T] for t in range(T)])
        s_new = 0.5 * np.einsum('nk,tkl,nl->nt', C, vsm, C)
        ds = np.abs(s_new - s).max()
        s = s_new
        if np.abs(step).max() <= tol * (1 + np.abs(m).max()) and ds <= tol * (1 + np.abs(s).max()):
            break
    lam = np.exp(C @ m + d[:, None] + s)
    D, grad, cov = dual_struct(lam, y, C, d, K, Kinv)
    mean = -np.einsum('kts,ks->kt', K, C.T @ (lam - y))
    vsmGP, vsm = slice_cov(cov, q, T)
    return {'lam': lam, 'mean': mean, 'cov': cov, 'vsm': vsm, 'vsmGP': vsmGP, 'D': D, 'grad': grad, 'iters': it + 1,
            'post_lik_term': nlp_struct(mean, y, C, d, Kinv)}


def dual_variational_struct(ys, params, T, binSize):
    """Structured, tightly converged dual variational E-step (list of per-trial dicts + the reference's scalars)."""
    C = np.asarray(params['C'], dtype=np.float64)
    d = np.ravel(np.asarray(params['d'], dtype=np.float64))
    K = make_K(params['tau'], T, binSize)
    Kinv = np.stack([np.linalg.inv(K[k]) for k in range(C.shape[1])])
    out = [dual_fixed_point_struct(y, C, d, K, Kinv) for y in ys]
    post_lik = -np.mean([o['post_lik_term'] for o in out])
    lower = np.mean([o['D'] for o in out])
    return out, post_lik, lower


def dense_dual_variational(experiment, params, optimizeLogLambda=False, prevOptimRes=None,
                           factr=None, pgtol=None):
    """funs/inference.py:259-432 (L-BFGS-B over lambda >= 1e-10 from 0.5, factr=1e7; or over rho)."""
    ys = _trial_list(experiment)
    N, T = ys[0].shape
    q = params['C'].shape[1]
    C_big, d_big = make_Cd_big(params, T)
    K_big, _ = make_K_big(params, experiment.trialDur, experiment.binSize)
    K_bigInv = np.linalg.inv(K_big)
    res = {'post_mean': [], 'post_cov': [], 'post_vsm': [], 'post_vsmGP': []}
    optim, tot_lik, tot_lb = [], 0.0, 0.0
    for r, y in enumerate(ys):
        ybar = y.reshape(N * T)
        args = (ybar, C_big, K_big, K_bigInv, d_big)
        kw = {}
        if pgtol is not None:
            kw['pgtol'] = pgtol
            kw['maxfun'] = kw['maxiter'] = 200000
        if not optimizeLogLambda:
            z0 = np.zeros(N * T) + 0.5 if prevOptimRes is None else prevOptimRes[r]
            out = sopt.fmin_l_bfgs_b(dense_dual, z0, fprime=dense_dual_grad, args=args,
                                     bounds=[(1e-10, None)] * (N * T),
                                     factr=1e7 if factr is None else factr, **kw)
            lam = out[0]
        else:
            z0 = np.zeros(N * T) if prevOptimRes is None else prevOptimRes[r]
            if factr is not None:
                kw['factr'] = factr
            out = sopt.fmin_l_bfgs_b(dense_dual_rho, z0, fprime=dense_dual_rho_grad, args=args, **kw)
            lam = np.exp(out[0])
        optim.append(out[0])
        mean = dense_vi_post_mean(K_big, C_big, ybar, lam)
        cov, _ = dense_vi_post_cov(K_bigInv, C_big, lam)
        tot_lik += dense_nlp(mean, ybar, C_big, d_big, K_bigInv)
        tot_lb += out[1]
        vsmGP, vsm = slice_cov(cov, q, T)
        res['post_mean'].append(mean.reshape(q, T))
        res['post_cov'].append(cov)
        res['post_vsm'].append(vsm)
        res['post_vsmGP'].append(vsmGP)
    R = len(ys)
    return res, -tot_lik / R, tot_lb / R, optim


def count_laplace_evals(ys, params, T, binSize):
    """Mean number of objective / gradient / Hessian evaluations scipy Newton-CG makes per trial in the reference's
    E-step (funs/inference.py:119-126, cold start, default options).  bench.py multiplies them with the cost of one
    evaluation at a shape where a whole reference E-step would take minutes per trial (labelled as an extrapolation)."""
    N = ys[0].shape[0]
    q = params['C'].shape[1]
    C_big, d_big = make_Cd_big(params, T)
    K_big, _ = make_K_big(params, T * binSize, binSize)
    K_bigInv = np.linalg.inv(K_big)
    tot = {"nfev": 0, "njev": 0, "nhev": 0}
    for y in ys:
        out = sopt.minimize(dense_nlp, np.zeros(q * T), args=(y.reshape(N * T), C_big, d_big, K_bigInv), method='Newton-CG',
                            jac=dense_nlp_grad, hess=dense_nlp_hess, options={'disp': False, 'maxiter': 10000})
        for k in tot:
            tot[k] += int(getattr(out, k))
    return {k: int(round(v / len(ys))) for k, v in tot.items()}


def count_dual_evals(ys, params, T, binSize):
    """Mean number of dual function(+gradient) evaluations L-BFGS-B makes per trial in the reference's variational E-step
    (funs/inference.py:316-324, cold start lambda = 0.5, factr = 1e7)."""
    N = ys[0].shape[0]
    C_big, d_big = make_Cd_big(params, T)
    K_big, _ = make_K_big(params, T * binSize, binSize)
    K_bigInv = np.linalg.inv(K_big)
    calls = 0
    for y in ys:
        out = sopt.fmin_l_bfgs_b(dense_dual, np.zeros(N * T) + 0.5, fprime=dense_dual_grad,
                                 args=(y.reshape(N * T), C_big, K_big, K_bigInv, d_big),
                                 bounds=[(1e-10, None)] * (N * T), factr=1e7)
        calls += int(out[2]['funcalls'])
    return {"funcalls": int(round(calls / len(ys)))}


# ----------------------------------------------------------------------------
# M-step: observation parameters C, d
# ----------------------------------------------------------------------------
def mstep_obs_cost(vecCd, xdim, ydim, ys, infRes):
    """funs/learning.py:20-49."""
    C, d = vec_to_Cd(vecCd, xdim, ydim)
    f = 0.0
    for y, m, vsm in zip(ys, infRes['post_mean'], infRes['post_vsm']):
        h = C @ m + d[:, None]
        rho = np.einsum('nk,tkl,nl->nt', C, vsm, C)
        f += (y * h - np.exp(h + 0.5 * rho)).sum()
    return -f / len(ys)


def mstep_obs_grad(vecCd, xdim, ydim, ys, infRes):
    """funs/learning.py:51-91: dC_n = sum_t (y-yhat) m_t - sum_t yhat V_t c_n ; dd_n = sum_t (y-yhat)."""
    C, d = vec_to_Cd(vecCd, xdim, ydim)
    dC = np.zeros_like(C)
    dd = np.zeros(ydim)
    for y, m, vsm in zip(ys, infRes['post_mean'], infRes['post_vsm']):
        h = C @ m + d[:, None]
        Vc = np.einsum('tkl,nl->ntk', vsm, C)
        yhat = np.exp(h + 0.5 * np.einsum('ntk,nk->nt', Vc, C))
        dC += (y - yhat) @ m.T - np.einsum('nt,ntk->nk', yhat, Vc)
        dd += (y - yhat).sum(1)
    return -Cd_to_vec(dC, dd) / len(ys)


def mstep_obs_cost_prior(vecCd, oldVec, xdim, ydim, ys, infRes, invPriorCov):
    """funs/learning.py:445-486: base cost - 0.5 D^T Lambda D (Lambda negative definite)."""
    dv = vecCd - oldVec
    return mstep_obs_cost(vecCd, xdim, ydim, ys, infRes) - 0.5 * dv @ (invPriorCov @ dv)


def mstep_obs_grad_prior(vecCd, oldVec, xdim, ydim, ys, infRes, invPriorCov):
    """funs/learning.py:488-534."""
    return mstep_obs_grad(vecCd, xdim, ydim, ys, infRes) - invPriorCov @ (vecCd - oldVec)


def learn_Cd(params, infRes, ys, method='TNC', options=None):
    """funs/learning.py:93-141."""
    N, q = params['C'].shape
    v0 = Cd_to_vec(params['C'], params['d'])
    opts = {'disp': False}
    opts.update(options or {})
    out = sopt.minimize(mstep_obs_cost, v0, args=(q, N, ys, infRes), jac=mstep_obs_grad,
                        method=method, options=opts)
    C, d = vec_to_Cd(out.x, q, N)
    return np.array(C), np.array(d), out.fun


def obs_stats_struct(C, d, ys, means, vsms):
    """Per-neuron cost, gradient and Hessian of R * MStepObservationCost (un-normalised sums over
    trials and bins) in the variables theta_n = (c_n, d_n): with u = m_t + V_t c_n,
      f_n = sum (yhat - y h), grad_c = sum (yhat u - y m), grad_d = sum (yhat - y),
      H_cc = sum yhat (u u^T + V_t), H_cd = sum yhat u, H_dd = sum yhat."""
    N, q = C.shape
    f = np.zeros(N)
    g = np.zeros((N, q + 1))
    H = np.zeros((N, q + 1, q + 1))
    for y, m, V in zip(ys, means, vsms):
        h = C @ m + d[:, None]
        Vc = np.einsum('tkl,nl->ntk', V, C)
        yhat = np.exp(h + 0.5 * np.einsum('ntk,nk->nt', Vc, C))
        u = Vc + m.T[None, :, :]
        f += (yhat - y * h).sum(1)
        g[:, :q] += np.einsum('nt,ntk->nk', yhat, u) - y @ m.T
        g[:, q] += (yhat - y).sum(1)
        H[:, :q, :q] += np.einsum('nt,ntk,ntl->nkl', yhat, u, u) + np.einsum('nt,tkl->nkl', yhat, V)
        H[:, :q, q] += np.einsum('nt,ntk->nk', yhat, u)
        H[:, q, q] += yhat.sum(1)
    H[:, q, :q] = H[:, :q, q]
    return f, g, H


def learn_Cd_newton(params, ys, means, vsms, prior_weight=0.0, tol=1e-13, maxit=200):
    """Exact per-neuron damped Newton on MStepObservationCost (+ 0.5*prior_weight*|theta-theta_old|^2,
    the 'useDiag' proximal term of funs/learning.py:580-581 with prior_weight = 1/s^2).
    The objective is separable over neurons and convex, so this is the fixed point every
    scipy method in funs/learning.py:124-130 / :605-622 approaches."""
    C = np.array(params['C'], dtype=np.float64)
    d = np.ravel(np.array(params['d'], dtype=np.float64))
    N, q = C.shape
    R = len(ys)
    th0 = np.concatenate([C, d[:, None]], axis=1)
    th = th0.copy()

    def evaluate(theta):
        with np.errstate(over='ignore', invalid='ignore'):
            f, g, H = obs_stats_struct(theta[:, :q], theta[:, q], ys, means, vsms)
        dt = theta - th0
        f = f / R + 0.5 * prior_weight * (dt ** 2).sum(1)
        g = g / R + prior_weight * dt
        H = H / R + prior_weight * np.eye(q + 1)[None]
        return f, g, H

    f, g, H = evaluate(th)
    for it in range(maxit):
        step = -np.linalg.solve(H, g[:, :, None])[:, :, 0]
        slope = (g * step).sum(1)
        a = np.ones(N)
        for _ in range(60):
            fn, gn, Hn = evaluate(th + a[:, None] * step)
            bad = ~(np.isfinite(fn) & (fn <= f + 1e-4 * a * slope + 1e-15 * np.abs(f)))
            if not bad.any():
                break
            a[bad] *= 0.5
        th = th + a[:, None] * step
        f, g, H = fn, gn, Hn
        if np.abs(a[:, None] * step).max() <= tol * (1.0 + np.abs(th).max()):
            break
    return th[:, :q].copy(), th[:, q].copy(), float(f.sum())


# ----------------------------------------------------------------------------
# M-step: GP timescales
# ----------------------------------------------------------------------------
def make_precomp(infRes):
    """funs/learning.py:145-173."""
    q, T = infRes['post_mean'][0].shape
    R = len(infRes['post_mean'])
    idx = np.arange(T)
    difSq = (idx[:, None] - idx[None, :]) ** 2
    out = []
    for k in range(q):
        P = np.zeros((T, T))
        for r in range(R):
            mk = infRes['post_mean'][r][k]
            P = P + infRes['post_vsmGP'][r][:, :, k] + np.outer(mk, mk)
        out.append({'T': T, 'difSq': difSq, 'numTrials': R, 'PautoSum': P})
    return out


def tau_cost(p, precomp, epsNoise=EPS_NOISE):
    """funs/learning.py:175-214: 0.5 R logdet K + 0.5 tr(K^-1 PautoSum), gamma = exp(p)."""
    p = float(np.ravel(p)[0])
    T = precomp['T']
    K = (1 - epsNoise) * np.exp(-np.exp(p) / 2 * precomp['difSq']) + epsNoise * np.eye(T)
    Kinv = np.linalg.inv(K)
    sign, ld = np.linalg.slogdet(K)
    return 0.5 * precomp['numTrials'] * sign * ld + 0.5 * (precomp['PautoSum'] * Kinv).sum()


def tau_cost_grad(p, precomp, epsNoise=EPS_NOISE):
    """funs/learning.py:216-255 without the half-matrix trick (identical value for even T):
    -dE/dgamma * gamma, dE/dgamma = -0.5 R tr(K^-1 dK) + 0.5 tr(K^-1 dK K^-1 P)."""
    p = float(np.ravel(p)[0])
    T = precomp['T']
    temp = (1 - epsNoise) * np.exp(-np.exp(p) / 2 * precomp['difSq'])
    K = temp + epsNoise * np.eye(T)
    dK = -0.5 * temp * precomp['difSq']
    Kinv = np.linalg.inv(K)
    KiM = Kinv @ dK
    dE = -0.5 * precomp['numTrials'] * np.trace(KiM) + 0.5 * ((KiM @ Kinv) * precomp['PautoSum'].T).sum()
    return np.array([-dE * np.exp(p)])


def tau_cost_prior(p, precomp, binSize, oldTau, step, epsNoise=EPS_NOISE):
    """funs/learning.py:681-724."""
    tau = binSize / 1000 * np.exp(-0.5 * float(np.ravel(p)[0]))
    return tau_cost(p, precomp, epsNoise) + 0.5 * (tau - oldTau) ** 2 / step ** 2


def tau_cost_prior_grad(p, precomp, binSize, oldTau, step, epsNoise=EPS_NOISE):
    """funs/learning.py:726-769. NOTE: the reference adds d(reg)/d(tau) to a d/dp gradient
    (no chain-rule factor) — reproduced as written."""
    tau = binSize / 1000 * np.exp(-0.5 * float(np.ravel(p)[0]))
    return tau_cost_grad(p, precomp, epsNoise) + (tau - oldTau) / step ** 2


def _polish_root(fun_grad, p, width=1e-3):
    """Zero of a scalar gradient near p: bracket + Brent to the last ulp.  scipy's BFGS ends with 'precision loss' in its
    line search up to |g| ~ 1e-5 away from the stationary point on some hosts (measured: p off by 1e-7 on the GPU boxes'
    CPUs, 1e-14 elsewhere, same inputs); the float64 root itself is the exact stationary point to ~1e-13 in p
    (tests/test_tau_search_host.py, 40-digit arithmetic).  Used only in tight (fixed-point parity) mode."""
    g = lambda v: float(np.ravel(fun_grad(v))[0])
    a, b = p - width, p + width
    ga, gb = g(a), g(b)
    for _ in range(8):
        if ga * gb < 0:
            break
        width *= 4
        a, b = p - width, p + width
        ga, gb = g(a), g(b)
    if ga * gb >= 0:
        return p
    return sopt.brentq(g, a, b, xtol=1e-15, rtol=1e-15)


def learn_tau(params, infRes, binSize, gtol=1e-8, polish=None):
    """funs/learning.py:257-293 (scipy default method = BFGS, gtol=1e-8).  polish (default: when gtol < 1e-8, i.e. the
    tight mode of the parity tests): refine BFGS's end point to the zero of the same gradient function."""
    q = infRes['post_mean'][0].shape[0]
    oldTau = np.ravel(params['tau']) * 1000 / binSize
    pre = make_precomp(infRes)
    new = np.zeros(q)
    details = []
    polish = (gtol < 1e-8) if polish is None else polish
    for k in range(q):
        p0 = np.log(1 / oldTau[k] ** 2)
        out = sopt.minimize(tau_cost, p0, args=(pre[k], EPS_NOISE), jac=tau_cost_grad,
                            options={'disp': False, 'gtol': gtol})
        if polish:
            out.x_bfgs = out.x.copy()
            out.x = np.array([_polish_root(lambda v: tau_cost_grad(v, pre[k]), float(out.x[0]))])
        details.append(out)
        new[k] = (1 / np.exp(out.x[0])) ** 0.5
    return new * binSize / 1000, details


def learn_tau_prior(params, infRes, binSize, step, method='TNC', gtol=1e-10):
    """funs/learning.py:771-830."""
    q = infRes['post_mean'][0].shape[0]
    tau_old = np.ravel(params['tau'])
    oldTau = tau_old * 1000 / binSize
    pre = make_precomp(infRes)
    new = np.zeros(q)
    details = []
    for k in range(q):
        p0 = np.log(1 / oldTau[k] ** 2)
        out = sopt.minimize(tau_cost_prior, p0, args=(pre[k], binSize, tau_old[k], step),
                            jac=tau_cost_prior_grad, method=method,
                            options={'disp': False, 'gtol': gtol})
        details.append(out)
        new[k] = (1 / np.exp(np.ravel(out.x)[0])) ** 0.5
    return new * binSize / 1000, details


# ----------------------------------------------------------------------------
# EM drivers
# ----------------------------------------------------------------------------
def update_params(params, infRes, experiment, CdOptimMethod='TNC', cd_options=None, tau_gtol=1e-8):
    """funs/learning.py:295-309."""
    ys = _trial_list(experiment)
    C, d, cost = learn_Cd(params, infRes, ys, CdOptimMethod, cd_options)
    tau, det = learn_tau(params, infRes, experiment.binSize, tau_gtol)
    return {'C': C, 'd': d, 'tau': tau}, {'Cd': cost, 'tau': det}


def batch_em_dense(experiment, initParams, maxEMiter, inferenceMethod='laplace', CdOptimMethod='TNC',
                   optimLogLamb=False):
    """The reference's batch EM loop, funs/engine.py:180-239 (dense port; CPU-baseline workload)."""
    params = copy.deepcopy(initParams)
    seq, lik, lb, t_inf, t_learn = [copy.deepcopy(params)], [], [], [], []
    import time
    prev = None
    for it in range(maxEMiter):
        t0 = time.time()
        if inferenceMethod == 'laplace':
            infRes, nll, prev = dense_laplace(experiment, params, prev)
        else:
            infRes, nll, vlb, prev = dense_dual_variational(experiment, params, optimLogLamb, prev)
            lb.append(vlb)
        lik.append(nll)
        t_inf.append(time.time() - t0)
        t0 = time.time()
        params, _ = update_params(params, infRes, experiment, CdOptimMethod)
        t_learn.append(time.time() - t0)
        seq.append(copy.deepcopy(params))
    return {'paramSeq': seq, 'posteriorLikelihood': lik, 'variationalLowerBound': lb,
            'inferenceTime': t_inf, 'learningTime': t_learn, 'infRes': infRes}


def em_step_struct(ys, params, T, binSize, x0s=None, tau_gtol=1e-11):
    """One tightly converged Laplace EM iteration (structured): the fixed-point target used for parity
    at sizes where the dense port is too slow. Returns (newParams, post_lik, modes, infRes)."""
    infRes, lik, optim, _ = laplace_struct(ys, params, T, binSize, x0s, want_cov=False)
    C, d, _ = learn_Cd_newton(params, ys, infRes['post_mean'], infRes['post_vsm'])
    tau, _ = learn_tau(params, infRes, binSize, gtol=tau_gtol)
    return {'C': C, 'd': d, 'tau': tau}, lik, optim, infRes


# ----------------------------------------------------------------------------
# Synthetic data of the named shapes (same distributions as funs/util.py:707-750, but sampled
# per latent through a Cholesky factor instead of an SVD of K_big: not stream-identical to the
# reference generator, used where only the shape/statistics matter)
# ----------------------------------------------------------------------------
class Experiment:
    """Duck-typed stand-in for funs/util.py:621 `dataset` (attributes read by the hot path:
    funs/engine.py:131-136, funs/util.py:463-472)."""

    def __init__(self, data, trialDur, binSize, params=None):
        self.data = data
        self.trialDur = trialDur
        self.binSize = binSize
        self.T = int(trialDur / binSize)
        self.numTrials = len(data)
        self.ydim = data[0]['Y'].shape[0]
        if params is not None:
            self.params = params
            self.xdim = params['C'].shape[1]


def synthetic_experiment(seed, xdim, ydim, numTrials, T, binSize=10, dOffset=-1.0, tau=None):
    rng = np.random.RandomState(seed)
    C = rng.rand(ydim, xdim) - 0.5
    d = rng.rand(ydim) * (-2) + dOffset
    tau = np.linspace(0.05, 0.3, xdim) if tau is None else np.asarray(tau, dtype=np.float64)
    K = make_K(tau, T, binSize)
    Lk = np.stack([np.linalg.cholesky(K[k]) for k in range(xdim)])
    data = []
    for _ in range(numTrials):
        X = np.einsum('kts,ks->kt', Lk, rng.randn(xdim, T))
        Y = rng.poisson(np.exp(C @ X + d[:, None])).astype(np.float64)
        data.append({'X': X, 'Y': Y})
    return Experiment(data, T * binSize, binSize, {'C': C, 'd': d, 'tau': tau.copy()})


def init_params_ppca(xdim, ydim, experiment, tau=None, seed=123):
    """Poisson-PCA initialisation, funs/util.py:505-558 (tau drawn as :557 unless given)."""
    ys = _trial_list(experiment)
    spikes = np.concatenate(ys, axis=1)
    meanY = spikes.mean(1) + 1e-10
    covY = np.cov(spikes)
    lamb = np.log(np.abs(covY + np.outer(meanY, meanY) - np.diag(meanY))) - np.log(np.outer(meanY, meanY))
    evals, evecs = np.linalg.eig(lamb)
    order = np.argsort(evals)[::-1]
    evecs = np.real(evecs[:, order][:, :xdim])
    if tau is None:
        tau = np.random.RandomState(seed).rand(xdim) * 0.5 + 0.1
    return {'C': np.ascontiguousarray(evecs), 'd': np.log(meanY), 'tau': np.asarray(tau, dtype=np.float64)}
