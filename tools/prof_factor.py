"""Profiling driver: one batched posterior factorisation + triangular inverse + slices at the bench shape.
Used under ncu (launch list / --set full); not a benchmark."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from poisson_gpfa_b200 import core, kernels as kn, _lib

R = int(os.environ.get("R", "1024"))
w = dict(bench.WORKLOAD); w["R"] = R
ex, ip = bench.make_data(w)
Y = _lib.dev_f64(np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data]))
p = core.DeviceParams(ip['C'], ip['d'], ip['tau'], w["T"], w["binSize"])
x = torch.zeros(R, w["q"], w["T"], dtype=torch.float64, device="cuda")
f, g, W = kn.laplace_eval(x, Y, p.C, p.d, p.Kinv)
bufs = kn.tile_buffers(R, w["q"] * w["T"], True)
torch.cuda.synchronize()
reps = int(os.environ.get("REPS", "1"))
for _ in range(reps):
    L, D, ZT, info = kn.potrf_posterior(p.Kinv, W, bufs=bufs)
    dx = kn.potrs(L, D, g.reshape(R, -1), scale=-1.0)
    kn.trtri(L, D, ZT, w["q"] * w["T"])
    vsm, vsmGP = kn.cov_slices(ZT, w["q"], w["T"])
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); kn.potrf_posterior(p.Kinv, W, bufs=bufs); e1.record(); torch.cuda.synchronize()
print("potrf_posterior (sequential schedule) ms:", e0.elapsed_time(e1), "TF:", R * 1600 ** 3 / 3 / e0.elapsed_time(e1) * 1e-9)
