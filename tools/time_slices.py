"""Time of the covariance-slice family (PautoSum product / post_vsmGP GEMM) of warm E-steps, alone on the GPU."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from poisson_gpfa_b200 import core, _lib
w = dict(bench.WORKLOAD); w["R"] = int(os.environ.get("R", "1024"))
ex, ip = bench.make_data(w)
Y = _lib.dev_f64(np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data]))
trials = core.DeviceTrials(Y, w["binSize"])
tau = np.asarray(ip['tau']) * float(os.environ.get("TAUSCALE", "1.0"))
p = core.DeviceParams(ip['C'], ip['d'], tau, w["T"], w["binSize"])
vs = os.environ.get("VSMGP", "0") == "1"
est = trials.estep_laplace(p, want_vsmGP=vs, want_pautosum=not vs)
h = _lib.handle()
_lib.call("pgpfa_set_profiling", h, 1)
for _ in range(5):
    est = trials.estep_laplace(p, x0=est.x, want_vsmGP=vs, want_pautosum=not vs)
torch.cuda.synchronize()
ms, work, cnt = bench.get_profile(_lib, h)
print(json.dumps({"lib": os.environ.get("PGPFA_LIB", "default"), "vsmGP": vs, "r": est.stats["lowrank_r"], "slices_ms": ms[4] / max(cnt[4], 1),
                  "tflops": work[4] / (ms[4] * 1e-3) / 1e12 if ms[4] else None, "lowrank_other_ms": ms[6] / 5, "factor_ms": ms[0] / 5}))
