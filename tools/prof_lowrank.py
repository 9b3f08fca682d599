"""Profiling driver: a few Laplace E-steps through the low-rank posterior pass at the bench shape (under ncu)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from poisson_gpfa_b200 import core, _lib

R = int(os.environ.get("R", "1024"))
w = dict(bench.WORKLOAD); w["R"] = R
ex, ip = bench.make_data(w)
Y = _lib.dev_f64(np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data]))
trials = core.DeviceTrials(Y, w["binSize"])
p = core.DeviceParams(ip['C'], ip['d'], ip['tau'], w["T"], w["binSize"])
if os.environ.get("MODE", "estep") == "estep":
    est = trials.estep_laplace(p, want_vsmGP=False, want_pautosum=True)
    for _ in range(int(os.environ.get("REPS", "1"))):
        est = trials.estep_laplace(p, x0=est.x, want_vsmGP=os.environ.get("VSMGP", "0") == "1", want_pautosum=True)
else:                                   # whole EM iterations (M-step kernels included)
    x0 = None
    for _ in range(1 + int(os.environ.get("REPS", "1"))):
        p, est, lik, info = trials.em_step(p, x0=x0)
        x0 = est.x
torch.cuda.synchronize()
print("ranks", p.lowrank[2], est.stats)
