"""Parity table on the GPU: deviations of one Laplace EM iteration from the exact-Newton oracle for two E-step
tolerances (which side of the tau deviation of round 1 was off: the covariance at the pre-polish point).
  python tools/parity_report.py > gpurun_out/parity_report.json"""
import copy
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pgpfa_oracle as po  # noqa: E402  (checker only)
from poisson_gpfa_b200 import core, _lib  # noqa: E402


def rel(a, b):
    a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
    b = np.asarray(b)
    return float(np.abs(a - b).max() / np.abs(b).max())


out = []
for (q, N, T, R, seed) in [(3, 12, 64, 6, 4), (2, 20, 50, 5, 7), (8, 100, 200, 4, 1), (10, 30, 250, 2, 3), (3, 10, 129, 1, 232)]:
    ex = po.synthetic_experiment(seed, q, N, R, T, binSize=10, dOffset=0.0 if N < 50 else -1.0)
    rng = np.random.RandomState(0)
    params = {'C': ex.params['C'] + 0.05 * rng.randn(N, q), 'd': ex.params['d'] + 0.05 * rng.randn(N),
              'tau': ex.params['tau'] * 1.2}
    ys = [np.asarray(t['Y'], dtype=np.float64) for t in ex.data]
    t0 = time.time()
    p_o, lik_o, _, ir = po.em_step_struct(ys, copy.deepcopy(params), T, 10)
    t_or = time.time() - t0
    for tol in (1e-8, 1e-10):
        for lowrank in ("1", "0"):
            os.environ["PGPFA_LOWRANK"] = lowrank
            trials = core.DeviceTrials(_lib.dev_f64(np.stack(ys)), 10)
            p = core.DeviceParams(params['C'], params['d'], params['tau'], T, 10)
            newp, est, lik, info = trials.em_step(p, tol=tol)
            row = {"shape": [q, N, T, R], "tol": tol, "lowrank": lowrank, "r": est.stats.get("lowrank_r"),
                   "x": rel(est.x, np.stack(ir['post_mean'])), "vsm": rel(est.vsm, np.stack(ir['post_vsm'])),
                   "vsmGP": rel(est.vsmGP, np.stack([v.transpose(2, 0, 1) for v in ir['post_vsmGP']])),
                   "lik": abs(lik - lik_o) / abs(lik_o), "C": rel(newp.C, p_o['C']), "d": rel(newp.d, p_o['d']),
                   "tau": rel(newp.tau, p_o['tau']), "newton": est.stats["pcg_newton_iters"], "cg": est.stats["pcg_iters"],
                   "cd_iters": info["cd_iters"], "tau_evals": info["tau_evals"], "oracle_s": round(t_or, 1)}
            out.append(row)
            print(json.dumps(row), flush=True)
