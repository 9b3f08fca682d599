"""Register-resident SPD inverse (gpprior.cu: spd_sweep_kernel) against numpy and against the tile-kernel chain; timing."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from poisson_gpfa_b200 import kernels as kn, _lib

def mk(batch, n, tau, eps=1e-3):
    i = np.arange(n)
    out = []
    for b in range(batch):
        K = (1 - eps) * np.exp(-0.5 * (i[:, None] - i[None, :]) ** 2 / (tau * (1 + 0.1 * b)) ** 2) + eps * np.eye(n)
        out.append(K)
    return np.stack(out)

for n, batch, tau in [(200, 72, 8.0), (200, 8, 20.0), (37, 5, 3.0), (208, 3, 5.0), (8, 2, 1.0), (1, 2, 1.0)]:
    A = mk(batch, n, tau)
    Ad = _lib.dev_f64(A)
    inv, ld, info = kn.spd_inverse(Ad)
    torch.cuda.synchronize()
    ref = np.linalg.inv(A); rld = np.linalg.slogdet(A)[1]
    e = np.abs(inv.cpu().numpy() - ref).max() / np.abs(ref).max()
    el = np.abs(ld.cpu().numpy() - rld).max() / np.abs(rld).max()
    resid = np.abs(A @ inv.cpu().numpy() - np.eye(n)).max()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3): kn.spd_inverse(Ad)
    ev0.record()
    for _ in range(20): kn.spd_inverse(Ad)
    ev1.record(); torch.cuda.synchronize()
    print("n=%d batch=%d  inv rel err %.2e  logdet rel err %.2e  |A inv - I| %.2e  info %s  %.1f us/call" %
          (n, batch, e, el, resid, info.cpu().numpy().tolist()[:4], ev0.elapsed_time(ev1) * 1e3 / 20), flush=True)
# a non-SPD matrix reports its first bad pivot
B = mk(2, 40, 3.0); B[1, 7, 7] = -1.0
inv, ld, info = kn.spd_inverse(_lib.dev_f64(B))
print("info for a bad pivot at 7:", info.cpu().numpy().tolist())
