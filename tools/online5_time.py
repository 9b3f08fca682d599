"""configs[4] with the reference's mini-batch of 5: per-step time and iteration counts (diagnosis of the small-batch path)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from poisson_gpfa_b200 import _lib, inference, learning, util
w = dict(bench.WORKLOAD_ONLINE); w["R"] = 2048
q, N, T, R = w["q"], w["N"], w["T"], w["R"]
ex = util.simulate_on_device(w["seed"], q, N, R, T, binSize=w["binSize"], dOffset=w["dOffset"], tau=np.linspace(0.05, 0.3, q))
rng = np.random.RandomState(5)
params = {'C': ex.params['C'] + 0.1 * rng.randn(N, q), 'd': ex.params['d'] + 0.1 * rng.randn(N), 'tau': ex.params['tau'] * 1.3}
np.random.seed(7)
inv_prior = np.eye(q * N + N)
B = int(os.environ.get("B", "5"))
for it in range(8):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    sub = util.subsampleTrials(ex, B)
    infRes, lik, _ = inference.laplace(sub, params)
    torch.cuda.synchronize(); t1 = time.perf_counter()
    step = 1.0 / (it + 1) ** 0.75
    params, det, inv_prior = learning.updateParamsWithPrior(params, infRes, sub, 'TNC', 'TNC', step, step, inv_prior, covOpts='useDiag')
    torch.cuda.synchronize(); t2 = time.perf_counter()
    st = infRes.device.stats
    print("it %d  E %.1f ms  M %.1f ms  newton %s pcg %s r %s fallback %s" % (it, (t1 - t0) * 1e3, (t2 - t1) * 1e3, st.get("pcg_newton_iters"), st.get("pcg_iters"), st.get("lowrank_r"), st.get("fallback_trials")), flush=True)
