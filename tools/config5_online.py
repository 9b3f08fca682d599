"""BASELINE.json configs[4]: stochastic mini-batch EM, 16384 trials, q=10 latents, N=200 neurons, T=250 bins
(qT = 2500 posterior systems), online 'diag' rule through engine.PPGPFAfit.  Prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from poisson_gpfa_b200 import engine, util, _lib

R = int(os.environ.get("R", "16384")); B = int(os.environ.get("BATCH", "512")); ITERS = int(os.environ.get("ITERS", "6"))
q, N, T = 10, 200, 250
t0 = time.time()
ex = util.simulate(1, q, N, R, T, binSize=10, dOffset=-1.0, tau=np.linspace(0.05, 0.3, q))
ex.Y_all = np.stack([t['Y'] for t in ex.data]).astype(np.float64)
gen_s = time.time() - t0
np.random.seed(123)
import contextlib, io
with contextlib.redirect_stdout(io.StringIO()):
    ip = util.initializeParams(q, N, ex)
ip = {k: np.real(np.asarray(v, dtype=np.complex128)).astype(np.float64) for k, v in ip.items()}
np.random.seed(7)
n0 = _lib.lib.pgpfa_launch_count()
torch.cuda.synchronize(); t0 = time.time()
fit = engine.PPGPFAfit(experiment=ex, initParams=ip, inferenceMethod='laplace', EMmode='Online', maxEMiter=ITERS,
                       batchSize=B, onlineParamUpdateMethod='diag', quiet=True)
torch.cuda.synchronize(); wall = time.time() - t0
print(json.dumps({"config": "configs[4] online 'diag' mini-batch EM", "trials": R, "q": q, "N": N, "T": T, "batchSize": B,
                  "iters": ITERS, "gen_s": gen_s, "wall_s": wall, "inference_s": fit.inferenceTime.tolist(),
                  "learning_s": fit.learningTime.tolist(), "post_lik": [float(v) for v in fit.posteriorLikelihood],
                  "tau": np.asarray(fit.optimParams['tau']).tolist(), "launches": int(_lib.lib.pgpfa_launch_count() - n0),
                  "minibatch_iters_per_s": (ITERS - 1) / float(fit.inferenceTime[1:].sum() + fit.learningTime[1:].sum())}))
