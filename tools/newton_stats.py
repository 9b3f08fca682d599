"""Diagnostic: per-trial Newton / chord step sequences at a steady-state EM iteration of the bench workload."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from poisson_gpfa_b200 import core, kernels as kn, _lib

w = dict(bench.WORKLOAD); w["R"] = int(os.environ.get("R", "256"))
ex, ip = bench.make_data(w)
Y = np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data])
trials = core.DeviceTrials(_lib.dev_f64(Y), w["binSize"])
T = w["T"]
params = core.DeviceParams(ip['C'], ip['d'], ip['tau'], T, w["binSize"])
x0 = None
hist = []
for it in range(int(os.environ.get("ITERS", "6"))):
    est = trials.estep_laplace(params, x0=x0, inexact_newton=False)
    C, d, cost, cd_it, _ = trials.mstep_cd(params, est)
    tau, det = trials.mstep_tau(params, trials.pautosum(est))
    newp = core.DeviceParams(C, d, tau, T, w["binSize"])
    hist.append(dict(it=it, newton=est.stats["max_newton_iters"], fact=est.stats["factorizations"],
                     dtau=float((newp.tau / params.tau - 1).abs().max()), dC=float((newp.C - params.C).abs().max())))
    # manual Newton from the warm start under the NEW params to log step sizes
    x = est.x.clone()
    L_old, D_old, _, _ = kn.potrf_posterior(params.Kinv, kn.laplace_eval(est.x, trials.y, params.C, params.d, params.Kinv)[2], want_zt=False)
    steps = []
    for k in range(5):
        f, g, W = kn.laplace_eval(x, trials.y, newp.C, newp.d, newp.Kinv)
        L, D, _, _ = kn.potrf_posterior(newp.Kinv, W, want_zt=False)
        dx = kn.potrs(L, D, g.reshape(x.shape[0], -1), scale=-1.0).reshape(x.shape)
        steps.append(dx.abs().amax(dim=(1, 2)).cpu().numpy())
        x = x + dx
    steps = np.stack(steps)
    hist[-1]["newton_steps_median"] = [float(np.median(s)) for s in steps]
    hist[-1]["newton_steps_max"] = [float(s.max()) for s in steps]
    # chord with the old factor
    x = est.x.clone(); cs = []
    for k in range(8):
        f, g, W = kn.laplace_eval(x, trials.y, newp.C, newp.d, newp.Kinv)
        dx = kn.potrs(L_old, D_old, g.reshape(x.shape[0], -1), scale=-1.0).reshape(x.shape)
        cs.append(dx.abs().amax(dim=(1, 2)).cpu().numpy()); x = x + dx
    cs = np.stack(cs)
    hist[-1]["chord_steps_median"] = [float(np.median(s)) for s in cs]
    hist[-1]["chord_rho_median"] = float(np.median(cs[3] / cs[2]))
    params, x0 = newp, est.x
print(json.dumps(hist, indent=1))
