"""Where does the wall time of a device-resident EM step go?  CUDA events at the phase boundaries of em_step (recorded on the
main stream) + host timestamps: GPU idle = wall - busy."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from poisson_gpfa_b200 import core, _lib, kernels as kn
w = dict(bench.WORKLOAD)
ex, ip = bench.make_data(w)
Y = _lib.dev_f64(np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data]))
trials = core.DeviceTrials(Y, w["binSize"])
p = core.DeviceParams(ip['C'], ip['d'], ip['tau'], w["T"], w["binSize"])
x0 = None
for _ in range(6):
    p, est, lik, info = trials.em_step(p, x0=x0); x0 = est.x
rows = []
for it in range(8):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    torch.cuda.synchronize(); t0 = time.perf_counter()
    ev[0].record()
    est = trials.estep_laplace(p, x0=x0, want_vsmGP=False, want_pautosum=True)
    t1 = time.perf_counter(); ev[1].record()
    cd = trials.mstep_cd_async(p, est)
    Psum = trials.pautosum(est)
    ts = trials.mstep_tau_async(p, Psum, n_blind=trials._tau_blind)
    ev[2].record()
    cd.join()
    q = p.q
    newp = core.DeviceParams(cd.th_cur[:, :q].contiguous(), cd.th_cur[:, q].contiguous(), ts.tau, trials.T, trials.binSize)
    pend = cd.pending() + ts.pending() + newp.pending()
    cd.join()
    ev[3].record()
    t2 = time.perf_counter()
    vals = core.read_packed(pend)
    t3 = time.perf_counter()
    ev[4].record()
    C, d, cost, cd_it, _ = cd.finish(vals[:3]); tau = ts.finish(vals[3:4]); newp.resolve(vals[4:])
    p, x0 = newp, est.x
    torch.cuda.synchronize(); t4 = time.perf_counter()
    g = lambda a, b: ev[a].elapsed_time(ev[b])
    rows.append({"wall_ms": (t4 - t0) * 1e3, "host_estep_call_ms": (t1 - t0) * 1e3, "host_enqueue_mstep_ms": (t2 - t1) * 1e3,
                 "host_wait_read_ms": (t3 - t2) * 1e3, "gpu_estep_ms": g(0, 1), "gpu_mstep_enq_span_ms": g(1, 3), "gpu_read_ms": g(3, 4)})
    print(json.dumps({k: round(v, 2) for k, v in rows[-1].items()}), flush=True)
