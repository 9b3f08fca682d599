"""Host-side cost of the end-to-end step (API with host buffers), piece by piece."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
from poisson_gpfa_b200 import inference, learning, core, _lib

w = dict(bench.WORKLOAD)
ex, ip = bench.make_data(w)
q, N, T, R = w["q"], w["N"], w["T"], w["R"]
n = q * T
Y_host = np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data])
Y_pin = torch.from_numpy(Y_host).pin_memory()
exp = bench.inference_experiment(Y_pin, w)
host_params = {k: v.copy() for k, v in ip.items()}
modes_host = None
rows = []
for i in range(8):
    t = {}
    torch.cuda.synchronize(); t0 = time.perf_counter()
    inference.upload_counts(exp); torch.cuda.synchronize(); t['upload_Y'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    prev = None
    if modes_host is not None:
        prev = np.zeros((R, n)); prev[:] = modes_host.reshape(R, n)
    t['prev_host'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    infRes, lik, optim = inference.laplace(exp, host_params, prevOptimRes=prev)
    torch.cuda.synchronize(); t['laplace'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    host_params, det = learning.updateParams(host_params, infRes, exp)
    torch.cuda.synchronize(); t['updateParams'] = time.perf_counter() - t0
    t0 = time.perf_counter()
    modes_host = optim.tensor.cpu().numpy()
    t['modes_d2h'] = time.perf_counter() - t0
    rows.append({k: round(v * 1e3, 2) for k, v in t.items()})
    print(json.dumps(rows[-1]), flush=True)
