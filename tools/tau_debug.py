"""Where does the device's timescale gradient lose digits?  (GPU diagnostic; checker = numpy oracle)"""
import os, sys, json
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pgpfa_oracle as po
from poisson_gpfa_b200 import kernels as kn, _lib, core

for (q, N, T, R, seed) in [(3, 12, 64, 6, 4), (3, 12, 200, 6, 4)]:
    ex = po.synthetic_experiment(seed, q, N, R, T, binSize=10, dOffset=0.0)
    rng = np.random.RandomState(0)
    params = {'C': ex.params['C'] + 0.05 * rng.randn(N, q), 'd': ex.params['d'] + 0.05 * rng.randn(N), 'tau': ex.params['tau'] * 1.2}
    ys = [np.asarray(t['Y'], dtype=np.float64) for t in ex.data]
    infRes, _, _, _ = po.laplace_struct(ys, params, T, 10, None, want_cov=False)
    pre = po.make_precomp(infRes)
    tau_o, det = po.learn_tau(params, infRes, 10, gtol=1e-11)
    p_o = np.array([dd.x[0] for dd in det])
    Psum = _lib.dev_f64(np.stack([pp['PautoSum'] for pp in pre]))
    c_dev, g_dev = kn.tau_eval(_lib.dev_f64(p_o), Psum, float(R), T)
    g_np = np.array([po.tau_cost_grad(p_o[k], pre[k])[0] for k in range(q)])
    c_np = np.array([po.tau_cost(p_o[k], pre[k]) for k in range(q)])
    gs = np.array([abs(po.tau_cost_grad(p_o[k] + 0.1, pre[k])[0]) for k in range(q)])
    print(json.dumps({"T": T, "g_dev": g_dev.cpu().tolist(), "g_np": g_np.tolist(), "g_scale": gs.tolist(),
                      "c_rel": (np.abs(c_dev.cpu().numpy() - c_np) / np.abs(c_np)).tolist()}))
    # pieces
    K, dK = kn.make_K_gamma(_lib.dev_f64(p_o), T)
    Kh, dKh = K.cpu().numpy(), dK.cpu().numpy()
    Kinv, logdet, info = kn.spd_inverse(K)
    Kih = Kinv.cpu().numpy()
    for k in range(q):
        temp = (1 - 0.001) * np.exp(-np.exp(p_o[k]) / 2 * pre[k]['difSq'])
        Kn = temp + 0.001 * np.eye(T); dKn = -0.5 * temp * pre[k]['difSq']
        Kin = np.linalg.inv(Kn)
        P = pre[k]['PautoSum']
        def traces(Ki, dKm):
            KiM = Ki @ dKm
            return np.trace(KiM), ((KiM @ Ki) * P.T).sum(), (Ki * P).sum()
        t_np = traces(Kin, dKn)
        t_dev_inv = traces(Kih[k], dKn)            # device inverse, numpy products
        t_devK = traces(np.linalg.inv(Kh[k]), dKh[k])   # device K/dK, numpy inverse
        print(json.dumps({"k": k, "K_rel": float(np.abs(Kh[k] - Kn).max()), "dK_rel": float(np.abs(dKh[k] - dKn).max() / np.abs(dKn).max()),
                          "Kinv_abs": float(np.abs(Kih[k] - Kin).max()), "Kinv_norm": float(np.abs(Kin).max()),
                          "sym_dev": float(np.abs(Kih[k] - Kih[k].T).max()),
                          "traces_np": [float(v) for v in t_np],
                          "dev_inverse_rel": [float(abs(a - b) / abs(b)) for a, b in zip(t_dev_inv, t_np)],
                          "dev_K_rel": [float(abs(a - b) / abs(b)) for a, b in zip(t_devK, t_np)]}))
