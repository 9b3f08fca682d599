"""Where the timescale M-step's time goes: device time of one pgpfa_tau_eval launch set vs the host-driven search."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from poisson_gpfa_b200 import core, _lib, kernels as kn

R = int(os.environ.get("R", "128"))
w = dict(bench.WORKLOAD); w["R"] = R
ex, ip = bench.make_data(w)
Y = _lib.dev_f64(np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data]))
trials = core.DeviceTrials(Y, w["binSize"])
T, q = w["T"], w["q"]
params = core.DeviceParams(ip['C'], ip['d'], ip['tau'], T, w["binSize"])
est = trials.estep_laplace(params)
Psum = trials.pautosum(est)
m = 9
P_rep = Psum.repeat(m, 1, 1).contiguous()
tau_rep = params.tau.repeat(m).contiguous()
p = _lib.dev_f64(np.log(1.0 / (np.tile(ip['tau'], m) * 1000 / w["binSize"]) ** 2))
ws = _lib.workspace(_lib.lib.pgpfa_tau_eval_workspace_bytes(q * m, T))
for _ in range(3):
    kn.tau_eval(p, P_rep, float(R), T, 1e-3, 0.0, tau_rep, w["binSize"], ws=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    kn.tau_eval(p, P_rep, float(R), T, 1e-3, 0.0, tau_rep, w["binSize"], ws=ws)
e1.record(); torch.cuda.synchronize()
dev_ms = e0.elapsed_time(e1) / 20
t0 = time.perf_counter()
for _ in range(20):
    c, g = kn.tau_eval(p, P_rep, float(R), T, 1e-3, 0.0, tau_rep, w["binSize"], ws=ws)
    torch.stack([c, g]).cpu()
sync_ms = (time.perf_counter() - t0) / 20 * 1e3
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    tau, det = trials.mstep_tau(params, Psum)
    torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
print(json.dumps({"tau_eval_device_ms": dev_ms, "tau_eval_with_readback_ms": sync_ms, "mstep_tau_ms": ts, "nfev": det["nfev"]}))
