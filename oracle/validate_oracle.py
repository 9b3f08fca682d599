"""Pins oracle/pgpfa_oracle.py against the UNMODIFIED reference (build container only).
Run: python oracle/validate_oracle.py   -> prints max relative deviations, exits non-zero on failure."""
import copy
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import pgpfa_oracle as po          # noqa: E402
from oracle import ref_harness as rh           # noqa: E402


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-300))


def main():
    ref = rh.load_reference()
    fails = []

    def check(name, val, tol):
        ok = val <= tol
        print("%-46s %.3e  (tol %.0e) %s" % (name, val, tol, "ok" if ok else "FAIL"))
        if not ok:
            fails.append(name)

    np.random.seed(123)
    with rh.quiet():
        ds = ref.util.dataset(seed=4711, xdim=3, ydim=12, numTrials=4, trialDur=400, binSize=10,
                              dOffset=0.0, fixTau=True, fixedTau=np.linspace(0.05, 0.2, 3))
        ip = ref.util.initializeParams(3, 12, ds)
    q, N, T = 3, 12, 40
    ys = [np.asarray(t['Y'], dtype=np.float64) for t in ds.data]

    # --- builders
    Kb_r, K_r = ref.util.makeK_big(copy.deepcopy(ip), ds.trialDur, ds.binSize)
    Kb_o, K_o = po.make_K_big(copy.deepcopy(ip), ds.trialDur, ds.binSize)
    check("makeK_big K", rel(K_o, K_r), 1e-15)
    check("makeK_big K_big", rel(Kb_o, Kb_r), 1e-15)
    Cb_r, db_r = ref.util.makeCd_big(ip, T)
    Cb_o, db_o = po.make_Cd_big(ip, T)
    check("makeCd_big", max(rel(Cb_o, Cb_r), rel(db_o, db_r)), 0)
    v = ref.util.CdtoVecCd(ip['C'], ip['d'])
    check("CdtoVecCd", rel(po.Cd_to_vec(ip['C'], ip['d']), v), 0)
    C2, d2 = po.vec_to_Cd(v, q, N)
    check("vecCdtoCd", max(rel(C2, ip['C']), rel(d2, ip['d'])), 0)

    # --- Laplace functions
    Kinv_big = np.linalg.inv(Kb_r)
    Kinv = np.stack([np.linalg.inv(K_r[k]) for k in range(q)])
    rng = np.random.RandomState(0)
    x = 0.3 * rng.randn(q * T)
    ybar = ys[0].reshape(-1)
    C, d = ip['C'], ip['d']
    f_r = ref.inference.negLogPosteriorUnNorm(x, ybar, Cb_r, db_r, Kinv_big, q, N)
    g_r = ref.inference.negLogPosteriorUnNorm_grad(x, ybar, Cb_r, db_r, Kinv_big, q, N)
    H_r = ref.inference.negLogPosteriorUnNorm_hess(x, ybar, Cb_r, db_r, Kinv_big, q, N)
    check("nlp dense", rel(po.dense_nlp(x, ybar, Cb_o, db_o, Kinv_big), f_r), 1e-14)
    check("nlp grad dense", rel(po.dense_nlp_grad(x, ybar, Cb_o, db_o, Kinv_big), g_r), 1e-13)
    check("nlp hess dense", rel(po.dense_nlp_hess(x, ybar, Cb_o, db_o, Kinv_big), H_r), 1e-13)
    X = x.reshape(q, T)
    check("nlp struct", rel(po.nlp_struct(X, ys[0], C, d, Kinv), f_r), 1e-12)
    check("nlp grad struct", rel(po.nlp_grad_struct(X, ys[0], C, d, Kinv).ravel(), g_r), 1e-11)
    check("nlp hess struct", rel(po.assemble_H(Kinv, po.nlp_W_struct(X, C, d)), H_r), 1e-11)

    # --- Laplace E-step: default tolerances (dense port == reference) and tight (struct == reference)
    with rh.quiet():
        ir_r, lik_r, opt_r = ref.inference.laplace(ds, copy.deepcopy(ip))
    ir_o, lik_o, opt_o = po.dense_laplace(ds, copy.deepcopy(ip))
    check("laplace default: post_mean", max(rel(a, b) for a, b in zip(ir_o['post_mean'], ir_r['post_mean'])), 1e-9)
    check("laplace default: post_cov", max(rel(a, b) for a, b in zip(ir_o['post_cov'], ir_r['post_cov'])), 1e-9)
    check("laplace default: post_lik", rel(lik_o, lik_r), 1e-12)
    with rh.tight_tolerances(), rh.quiet():
        ir_t, lik_t, opt_t = ref.inference.laplace(ds, copy.deepcopy(ip))
    ir_s, lik_s, opt_s, _ = po.laplace_struct(ys, copy.deepcopy(ip), T, ds.binSize)
    check("laplace tight: post_mean", max(rel(a, b) for a, b in zip(ir_s['post_mean'], ir_t['post_mean'])), 1e-9)
    check("laplace tight: post_cov", max(rel(a, b) for a, b in zip(ir_s['post_cov'], ir_t['post_cov'])), 1e-9)
    check("laplace tight: post_vsm", max(rel(a, b) for a, b in zip(ir_s['post_vsm'], ir_t['post_vsm'])), 1e-9)
    check("laplace tight: post_vsmGP", max(rel(a, b) for a, b in zip(ir_s['post_vsmGP'], ir_t['post_vsmGP'])), 1e-9)
    check("laplace tight: post_lik", rel(lik_s, lik_t), 1e-12)
    print("   (reference default vs tight: mean %.2e cov %.2e)" % (
        max(rel(a, b) for a, b in zip(ir_r['post_mean'], ir_t['post_mean'])),
        max(rel(a, b) for a, b in zip(ir_r['post_cov'], ir_t['post_cov']))))

    # --- M-step C,d
    vec = po.Cd_to_vec(C, d) + 0.01 * rng.randn(q * N + N)
    c_r = ref.learning.MStepObservationCost(vec, q, N, ds, ir_t)
    gq_r = ref.learning.MStepObservationCost_grad(vec, q, N, ds, ir_t)
    check("MStepObservationCost", rel(po.mstep_obs_cost(vec, q, N, ys, ir_t), c_r), 1e-13)
    check("MStepObservationCost_grad", rel(po.mstep_obs_grad(vec, q, N, ys, ir_t), gq_r), 1e-12)
    Cv, dv = po.vec_to_Cd(vec, q, N)
    f_n, g_n, H_n = po.obs_stats_struct(Cv, dv, ys, ir_t['post_mean'], ir_t['post_vsm'])
    check("obs_stats cost sum", rel(f_n.sum() / len(ys), c_r), 1e-13)
    check("obs_stats grad", rel(po.Cd_to_vec(g_n[:, :q], g_n[:, q]) / len(ys), gq_r), 1e-12)
    eps = 1e-6
    e = np.zeros(q + 1); e[1] = eps
    fp, gp, _ = po.obs_stats_struct(Cv + e[None, :q], dv + e[q], ys, ir_t['post_mean'], ir_t['post_vsm'])
    fm, gm, _ = po.obs_stats_struct(Cv - e[None, :q], dv - e[q], ys, ir_t['post_mean'], ir_t['post_vsm'])
    check("obs_stats hess (FD col 1)", rel((gp - gm) / (2 * eps), H_n[:, :, 1]), 1e-6)
    Lam = -np.eye(q * N + N) / 0.37 ** 2
    oldv = po.Cd_to_vec(C, d)
    cp_r = ref.learning.MStepObservationCostWithPrior(vec, ip, q, N, ds, ir_t, Lam)
    gp_r = ref.learning.MStepObservationCostWithPrior_grad(vec, ip, q, N, ds, ir_t, Lam)
    check("MStepObservationCostWithPrior", rel(po.mstep_obs_cost_prior(vec, oldv, q, N, ys, ir_t, Lam), cp_r), 1e-13)
    check("MStepObservationCostWithPrior_grad", rel(po.mstep_obs_grad_prior(vec, oldv, q, N, ys, ir_t, Lam), gp_r), 1e-12)
    with rh.tight_tolerances(), rh.quiet():
        C_r, d_r, cost_r = ref.learning.learnLTparams(copy.deepcopy(ip), ir_t, ds, 'TNC')
    C_n, d_n, cost_n = po.learn_Cd_newton(ip, ys, ir_t['post_mean'], ir_t['post_vsm'])
    # scipy's line-search optimisers stall at |grad| ~ 1e-7 on this cost (function-value noise), so the
    # reference cannot be driven closer than ~1e-7 to the optimum; the Newton fixed point is certified
    # with the REFERENCE's own gradient instead.
    check("learnLTparams tight vs Newton: C", rel(C_n, C_r), 1e-6)
    check("learnLTparams tight vs Newton: d", rel(d_n, d_r), 1e-6)
    check("reference gradient at Newton optimum", float(np.abs(ref.learning.MStepObservationCost_grad(
        po.Cd_to_vec(C_n, d_n), q, N, ds, ir_t)).max()), 1e-12)
    check("learnLTparams tight vs Newton: cost", rel(cost_n / len(ys) if False else cost_n, cost_r * 1.0), 1e-10)

    # --- M-step tau
    pre_r = ref.learning.makePrecomp(ir_t)
    pre_o = po.make_precomp(ir_t)
    check("makePrecomp PautoSum", max(rel(a['PautoSum'], b['PautoSum']) for a, b in zip(pre_o, pre_r)), 1e-14)
    p = np.log(1 / (ip['tau'][1] * 1000 / ds.binSize) ** 2) + 0.1
    check("MStepGPtimescaleCost", rel(po.tau_cost(p, pre_o[1]), ref.learning.MStepGPtimescaleCost(p, pre_r[1], 0.001)), 1e-12)
    check("MStepGPtimescaleCost_grad", rel(po.tau_cost_grad(p, pre_o[1]), ref.learning.MStepGPtimescaleCost_grad(p, pre_r[1], 0.001)), 1e-9)
    check("MStepGPtimescaleCostWithPrior", rel(po.tau_cost_prior(p, pre_o[1], ds.binSize, ip['tau'][1], 0.5),
                                                 ref.learning.MStepGPtimescaleCostWithPrior(p, pre_r[1], 0.001, ds.binSize, ip['tau'][1], 0.5)), 1e-12)
    check("MStepGPtimescaleCostWithPrior_grad", rel(po.tau_cost_prior_grad(p, pre_o[1], ds.binSize, ip['tau'][1], 0.5),
                                                      ref.learning.MStepGPtimescaleCostWithPrior_grad(p, pre_r[1], 0.001, ds.binSize, ip['tau'][1], 0.5)), 1e-9)
    with rh.quiet():
        tau_r, _ = ref.learning.learnGPparams(copy.deepcopy(ip), ir_t, ds)
    tau_o, _ = po.learn_tau(copy.deepcopy(ip), ir_t, ds.binSize)
    check("learnGPparams (default gtol)", rel(tau_o, tau_r), 1e-9)

    # --- VI functions (small)
    lam = np.exp(0.3 * rng.randn(N * T))
    D_r = ref.inference.dualProblem(lam, ybar, Cb_r, Kb_r, Kinv_big, db_r)
    Dg_r = ref.inference.dualProblem_grad(lam, ybar, Cb_r, Kb_r, Kinv_big, db_r)
    check("dualProblem dense", rel(po.dense_dual(lam, ybar, Cb_o, Kb_o, Kinv_big, db_o), D_r), 1e-12)
    check("dualProblem_grad dense", rel(po.dense_dual_grad(lam, ybar, Cb_o, Kb_o, Kinv_big, db_o), Dg_r), 1e-10)
    D_s, g_s, cov_s = po.dual_struct(lam.reshape(N, T), ys[0], C, d, K_r, Kinv)
    check("dualProblem struct", rel(D_s, D_r), 1e-11)
    check("dualProblem_grad struct", rel(g_s.ravel(), Dg_r), 1e-9)
    check("VIPostCov struct", rel(cov_s, ref.inference.VIPostCov(Kinv_big, Cb_r, lam)[0]), 1e-9)

    # --- full batch EM, 2 iterations, default tolerances: dense port == reference
    with rh.quiet():
        fit = ref.engine.PPGPFAfit(experiment=ds, initParams=copy.deepcopy(ip), inferenceMethod='laplace',
                                   EMmode='Batch', maxEMiter=2)
    out = po.batch_em_dense(ds, copy.deepcopy(ip), 2)
    check("batch EM x2 (default tol): C", rel(out['paramSeq'][-1]['C'], fit.optimParams['C']), 1e-5)
    check("batch EM x2 (default tol): tau", rel(out['paramSeq'][-1]['tau'], fit.optimParams['tau']), 1e-7)
    check("batch EM x2 (default tol): lik", rel(out['posteriorLikelihood'], fit.posteriorLikelihood), 1e-9)

    print("FAILED: %s" % fails if fails else "oracle pinned against the reference: all checks ok")
    return 1 if fails else 0


if __name__ == "__main__":
    sys.exit(main())
