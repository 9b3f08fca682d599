"""Wall-clock breakdown of one steady-state EM iteration (E-step / C,d / PautoSum / tau) at a given trial count."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from poisson_gpfa_b200 import core, _lib

R = int(os.environ.get("R", "128"))
w = dict(bench.WORKLOAD); w["R"] = R
ex, ip = bench.make_data(w)
Y = _lib.dev_f64(np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data]))
trials = core.DeviceTrials(Y, w["binSize"])
T = w["T"]
params = core.DeviceParams(ip['C'], ip['d'], ip['tau'], T, w["binSize"])
x0 = None
def tick():
    torch.cuda.synchronize(); return time.perf_counter()
rows = []
for it in range(8):
    t0 = tick(); est = trials.estep_laplace(params, x0=x0)
    t1 = tick(); lik = trials.post_lik(est)
    C, d, cost, cd_it, _ = trials.mstep_cd(params, est)
    t2 = tick(); Psum = trials.pautosum(est)
    t3 = tick(); tau, det = trials.mstep_tau(params, Psum)
    t4 = tick(); params = core.DeviceParams(C, d, tau, T, w["binSize"]); _ = params.Kinv
    t5 = tick(); x0 = est.x
    rows.append(dict(it=it, estep=(t1-t0)*1e3, cd=(t2-t1)*1e3, pauto=(t3-t2)*1e3, tau=(t4-t3)*1e3, kinv=(t5-t4)*1e3,
                     cd_it=cd_it, nfev=det['nfev'], newton=est.stats['max_newton_iters'], chord=est.stats['pcg_iters']))
print(json.dumps(rows[3:], indent=0))
