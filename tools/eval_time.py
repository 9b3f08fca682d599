"""Device time of laplace_eval / linesearch-sized launches at the headline shape (all trials active)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from poisson_gpfa_b200 import core, _lib, kernels as kn
w = dict(bench.WORKLOAD)
ex, ip = bench.make_data(w)
Y = _lib.dev_f64(np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data]))
trials = core.DeviceTrials(Y, w["binSize"])
p = core.DeviceParams(ip['C'], ip['d'], ip['tau'], w["T"], w["binSize"])
est = trials.estep_laplace(p, want_vsmGP=False, want_pautosum=True)
f, g, W = kn.laplace_eval(est.x, trials.y, p.C, p.d, p.Kinv)
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(20):
    f2, g2, W2 = kn.laplace_eval(est.x, trials.y, p.C, p.d, p.Kinv)
ev1.record(); torch.cuda.synchronize()
print("laplace_eval (prior apply + eval): %.1f us/call; f sum %.12e  |g| %.3e  W sum %.12e" %
      (ev0.elapsed_time(ev1) * 1e3 / 20, float(f2.sum()), float(g2.abs().max()), float(W2.sum())))
