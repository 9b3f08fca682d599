#!/usr/bin/env python
"""Headline benchmark: EM iterations / second of full-batch Laplace EM (E-step + M-step) on
1024 trials, q=8 latents, N=100 neurons, T=200 bins (BASELINE.json configs[2]), synthetic data.

  python bench.py [--gpus N --steps K --warmup W]            our arm (one process per GPU under torchrun)
  python bench.py --impl reference [--steps K --warmup W]    the reference algorithm on the host CPU

A step is one EM iteration (E-step + M-step) of one fit; warm-up iterations are the first W
iterations of the same fit (cold start included), the timed K iterations follow directly, i.e.
steady-state warm-started EM.  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOAD = dict(q=8, N=100, T=200, R=1024, binSize=10, dOffset=-1.0, seed=1)
METRIC = "EM iters/sec (Laplace E+M, 1024 trials q=8 T=200)"


def make_data(w, R=None):
    from poisson_gpfa_b200 import util
    R = w["R"] if R is None else R
    ex = util.simulate(w["seed"], w["q"], w["N"], R, w["T"], binSize=w["binSize"], dOffset=w["dOffset"])
    np.random.seed(123)
    import contextlib, io
    with contextlib.redirect_stdout(io.StringIO()):
        ip = util.initializeParams(w["q"], w["N"], ex)
    ip = {k: np.ascontiguousarray(np.real(v), dtype=np.float64) for k, v in ip.items()}
    return ex, ip


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.proc, self.lines, self.index = None, [], index
        self.t0 = self.t1 = None

    def start(self):
        """Started before the warm-up so that nvidia-smi is already sampling (every 50 ms) when the timed region begins;
        samples carry their arrival time and only those inside [begin(), end()] are reported."""
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append((time.perf_counter(), l)) for l in self.proc.stdout],
                             daemon=True).start()
        except Exception:
            self.proc = None

    def begin(self):
        self.t0 = time.perf_counter()

    def end(self):
        self.t1 = time.perf_counter()

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        t0 = self.t0 if self.t0 is not None else -1e300
        t1 = self.t1 if self.t1 is not None else 1e300
        inside = [l for (t, l) in self.lines if t0 <= t <= t1 + 0.03]
        lines = inside if inside else [l for (_, l) in self.lines]     # region shorter than one sampling period
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for l in lines:
            p = [s.strip() for s in l.split(",")]
            try:
                sm.append(float(p[0])); mx = float(p[1])
            except Exception:
                continue
            for nme, v in zip(names, p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "samples": len(sm),
                "samples_in_timed_region": len(inside), "reasons": sorted(reasons)}


def cpu_sample(w, R_cpu, warm, timed):
    """The reference algorithm (oracle dense port: same C_big / K_big formulation, same scipy optimisers and
    options as funs/inference.py + funs/learning.py, funs/engine.py:180-239 loop) on a bounded sample of the same
    workload: `warm` untimed + `timed` timed EM iterations of ONE fit (the same protocol as the GPU arm: the timed
    iterations are warm-started).  Returns seconds per timed EM iteration."""
    from oracle import pgpfa_oracle as po
    ex, ip = make_data(w, R_cpu)
    ex_o = po.Experiment([{'Y': np.asarray(t['Y'], dtype=np.float64)} for t in ex.data], ex.trialDur, ex.binSize)
    out = po.batch_em_dense(ex_o, ip, warm + timed)
    per_it = np.asarray(out['inferenceTime']) + np.asarray(out['learningTime'])
    return float(per_it[warm:].mean()), out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    w = dict(WORKLOAD)
    R_cpu = args.cpu_trials
    sec, _ = cpu_sample(w, R_cpu, args.warmup_ref, args.steps)
    value = 1.0 / (sec * w["R"] / R_cpu)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "EM iters/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup_ref, "ms_per_step": 1e3 / value, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[2]: synthetic q=8 N=100 T=200 R=1024 full-batch Laplace EM, steady-state "
                                   "(warm-started) iterations"},
            "cpu_baseline": {"value": value, "unit": "EM iters/s", "cores": os.cpu_count(), "kind": "port",
                             "sample": "%d warm-started EM iterations (dense C_big formulation, scipy Newton-CG/TNC/BFGS "
                                       "at the reference's options) on %d of 1024 trials, %.1f s each; per-trial cost is "
                                       "exactly linear in trials (serial loops funs/inference.py:94, "
                                       "funs/learning.py:39), extrapolated x%d" % (args.steps, R_cpu, sec, w["R"] // R_cpu)},
            "e2e": {"value": value, "unit": "EM iters/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def run_ours(args):
    import torch
    from poisson_gpfa_b200 import _lib, core, dist, inference, learning
    # keep stdout clean for the single JSON line: NCCL's banner goes to stderr
    if os.environ.get("NCCL_DEBUG", "").upper() not in ("INFO", "TRACE"):
        os.environ["NCCL_DEBUG"] = "WARN"
    sys.stdout.flush()
    saved_fd = os.dup(1)
    os.dup2(2, 1)
    try:
        red = dist.init_from_env()
        if red.world_size > 1:
            red.sum_scalar(1.0)            # forces communicator creation inside the redirected region
            torch.cuda.synchronize()
    finally:
        sys.stdout.flush()
        os.dup2(saved_fd, 1)
        os.close(saved_fd)
    rank, world = red.rank, red.world_size
    if world == 1:
        torch.cuda.set_device(0)
    dev_index = torch.cuda.current_device()
    w = dict(WORKLOAD)
    if args.trials:
        w["R"] = args.trials
    ex, ip = make_data(w)
    q, N, T, R = w["q"], w["N"], w["T"], w["R"]
    n = q * T

    Y_host = np.stack([np.asarray(t['Y'], dtype=np.float64) for t in ex.data])
    lo, hi = dist.shard_bounds(R, world, rank)
    trials = core.DeviceTrials(_lib.dev_f64(Y_host[lo:hi]), w["binSize"], red, R_total=R, offset=lo)
    h = _lib.handle()

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def em_iteration(params, x0):
        newp, est, lik, info = trials.em_step(params, x0=x0)
        return newp, est, lik, info["cd_iters"], info["tau_evals"]

    params = core.DeviceParams(ip['C'], ip['d'], ip['tau'], T, w["binSize"])
    sampler = ClockSampler(dev_index)
    if rank == 0:
        sampler.start()                       # already sampling when the timed region begins
    x0, liks = None, []
    for _ in range(args.warmup):
        params, est, lik, _, _ = em_iteration(params, x0)
        x0 = est.x
        liks.append(lik)

    # ---------------- timed region: K steady-state EM iterations, device-resident inputs
    _lib.call("pgpfa_set_profiling", h, 1)
    launches0 = _lib.lib.pgpfa_launch_count()
    syncs0 = _lib.host_sync_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    sampler.begin()
    e0.record()
    newton_its, facts, cd_its, tau_evals, chord_its, fallback = [], 0, [], [], [], []
    for _ in range(args.steps):
        params, est, lik, cd_it, nfev = em_iteration(params, x0)
        x0 = est.x
        liks.append(lik)
        newton_its.append(est.stats["max_newton_iters"]); facts += est.stats["factorizations"]
        cd_its.append(cd_it); tau_evals.append(nfev)
        chord_its.append(est.stats["pcg_newton_iters"]); fallback.append(est.stats["pcg_iters"])
    e1.record()
    barrier()
    sampler.end()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = _lib.lib.pgpfa_launch_count() - launches0
    host_syncs = _lib.host_sync_count() - syncs0
    prof_ms, prof_work, prof_cnt = (ctypes.c_double * 8)(), (ctypes.c_double * 8)(), (ctypes.c_longlong * 8)()
    _lib.call("pgpfa_get_profile", h, ctypes.cast(prof_ms, ctypes.c_void_p), ctypes.cast(prof_work, ctypes.c_void_p),
              ctypes.cast(prof_cnt, ctypes.c_void_p))
    _lib.call("pgpfa_set_profiling", h, 0)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t.item())
    value = args.steps / (ms * 1e-3)

    # ---------------- e2e: the same EM iteration through the public API with HOST buffers every step
    e2e_steps = max(1, min(args.e2e_steps, args.steps)) if not args.profile_mode else 0
    Y_pin = torch.from_numpy(Y_host).pin_memory()
    host_params = params.to_numpy_dict()
    modes_host = est.x.cpu().numpy()            # this rank's modes
    e2e_t = []
    h2d = d2h = 0
    exp = inference_experiment(Y_pin, w)
    E2E_WARMUP = 2                                         # untimed: allocate the API path's own buffers, settle the host
    for i in range(e2e_steps + E2E_WARMUP if e2e_steps else 0):
        barrier()
        t0 = time.perf_counter()
        inference.upload_counts(exp)                       # H2D of this step's inputs (pinned -> HBM)
        prev = None
        if modes_host is not None:                         # per-trial modes of the previous iteration, host side
            prev = np.zeros((R, n))
            prev[lo:hi] = modes_host.reshape(hi - lo, n)
        infRes, lik_e, optim = inference.laplace(exp, host_params, prevOptimRes=prev, reducer=red)
        host_params, det = learning.updateParams(host_params, infRes, exp)
        modes_host = optim.tensor.cpu().numpy()            # D2H of the step's results
        barrier()
        if i >= E2E_WARMUP:
            e2e_t.append(time.perf_counter() - t0)
        h2d = (hi - lo) * N * T * 8 + (hi - lo) * n * 8 + (N * q + N + q) * 8
        d2h = (hi - lo) * n * 8 + (N * q + N + q) * 8 + 8
    e2e_sec = float(np.mean(e2e_t)) if e2e_t else float("nan")
    print("e2e step times (s):", [round(t, 4) for t in e2e_t], file=sys.stderr)
    if world > 1:
        t = torch.tensor([e2e_sec], dtype=torch.float64, device="cuda")
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e_sec = float(t.item())

    if rank != 0:
        if world > 1:
            torch.distributed.barrier()
            torch.distributed.destroy_process_group()
        return
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "profiles", "r01_fp64_peaks.json")))
    except Exception:
        pass
    fp64_peak = float(peaks.get("fp64_roofline_peak_tflops", 35.4))
    lowrank_r = int(est.stats.get("lowrank_r", 0))
    other = {"solves": prof_ms[1] / args.steps, "eval_linesearch": prof_ms[2] / args.steps,
             "trtri": prof_ms[3] / args.steps, "cov_slices": prof_ms[4] / args.steps,
             "factor": prof_ms[0] / args.steps, "cg_preconditioner_setup": prof_ms[5] / args.steps,
             "lowrank_posterior_other": prof_ms[6] / args.steps}

    def traffic_of(fname, key):
        try:
            return json.load(open(os.path.join(ROOT, "profiles", fname)))[key] * (hi - lo) / 1024.0   # captured at 1024 trials
        except Exception:
            return None

    peak_src = ("measured cuBLAS DGEMM 8192^3 sustained on this pool (profiles/r01_fp64_peaks.json); "
                "MEASURED_PEAKS.json has no FP64 line")
    if lowrank_r > 0:
        # low-rank posterior pass (csrc/lowrank.cu): the dominant launch is the batched symmetric product
        # post_vsmGP[k] = eps diag(P) + Y_k Y_k^T (gemm_nt_kernel, q x trials problems of T x T x r)
        k_ms, k_flops, k_cnt = prof_ms[4], prof_work[4], prof_cnt[4]
        achieved = k_flops / (k_ms * 1e-3) / 1e12 if k_ms > 0 else 0.0
        roofline = {"bound": "tensor",
                    "kernel": "gemm_nt_kernel, the post_vsmGP launch (q*trials symmetric T x T x r products Y_k Y_k^T on DMMA.8x8x4); "
                              "r = %d is the rank of the prior factor; the C,d M-step runs concurrently on a second stream" % lowrank_r,
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                    "traffic": traffic_of("r01_syrk_traffic.json", "dram_bytes_per_launch"),
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of that launch, ncu --set full capture "
                                      "(profiles/r01_syrk_traffic.json), scaled to this rank's trial count",
                    "peak_source": peak_src,
                    "algorithmic_flops_per_launch": k_flops / max(k_cnt, 1), "launches": int(k_cnt),
                    "share_of_step": k_ms / ms, "other_ms_per_step": other,
                    "rxr_cholesky_tflops": prof_work[0] / (prof_ms[0] * 1e-3) / 1e12 if prof_ms[0] > 0 else None}
    else:
        fac_ms, fac_flops, fac_cnt = prof_ms[0], prof_work[0], prof_cnt[0]
        achieved = fac_flops / (fac_ms * 1e-3) / 1e12 if fac_ms > 0 else 0.0
        roofline = {"bound": "tensor", "kernel": "batched Cholesky call (chol_diag_kernel + chol_panel_kernel launches, DMMA.8x8x4); a launch = one factorisation of all trials of the rank",
                    "achieved": achieved, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved / fp64_peak,
                    "traffic": traffic_of("r01_factor_traffic.json", "dram_bytes_per_factorisation"),
                    "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum over the launches of one factorisation call, "
                                      "ncu launch list of this command (profiles/r01_factor_traffic.json), scaled to this rank's "
                                      "trial count",
                    "peak_source": peak_src,
                    "algorithmic_flops_per_launch": fac_flops / max(fac_cnt, 1), "launches": int(fac_cnt),
                    "share_of_step": fac_ms / ms, "other_ms_per_step": other,
                    "trtri_tflops": prof_work[3] / (prof_ms[3] * 1e-3) / 1e12 if prof_ms[3] > 0 else None,
                    "solve_gbs": prof_work[1] / (prof_ms[1] * 1e-3) / 1e9 if prof_ms[1] > 0 else None}
    cpu = None
    if not args.skip_cpu and not args.profile_mode and world == 1:
        sec, _ = cpu_sample(w, args.cpu_trials, 1, 1)
        v = 1.0 / (sec * R / args.cpu_trials)
        cpu = {"value": v, "unit": "EM iters/s", "cores": os.cpu_count(), "kind": "port",
               "sample": "second (warm-started) EM iteration of the dense reference formulation on %d of %d trials "
                         "(%.1f s), linear extrapolation in trials" % (args.cpu_trials, R, sec)}
    line = {"metric": METRIC, "value": value, "unit": "EM iters/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "configs[2]: synthetic q=8 N=100 T=200 R=%d full-batch Laplace EM, steady-state "
                                   "(warm-started) iterations" % R,
                       "posterior_pass": ("low-rank prior factor, r=%d of qT=%d" % (lowrank_r, n)) if lowrank_r else "dense tiled Cholesky",
                       "trials_per_gpu": hi - lo,
                       "l2": ("working set (Y, post_vsmGP, counts: %.1f GB/GPU) >> L2, no flush needed"
                              % ((hi - lo) * (n * lowrank_r + q * T * T + N * T) * 8 / 1e9)) if lowrank_r else
                             ("working set (factor tiles %.1f GB/GPU) >> L2, no flush needed" % ((hi - lo) * 2 * 10.65e6 / 1e9)),
                       "newton_tol": 1e-8},
            "clocks": clocks, "gpu_launches": int(launches),
            "e2e": {"value": 1.0 / e2e_sec, "unit": "EM iters/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h), "steps": e2e_steps,
                    "api": "inference.laplace + learning.updateParams with host numpy inputs/outputs each step"},
            "roofline": roofline, "cpu_baseline": cpu,
            "detail": {"newton_iters_per_step": newton_its, "trial_factorisations": facts, "cd_newton_iters": cd_its,
                       "tau_evals": tau_evals, "inexact_newton_iters_per_step": chord_its, "pcg_iters_per_step": fallback, "post_lik": liks[-3:], "allreduces": red.n_allreduce,
                       "host_syncs_per_step": host_syncs / args.steps}}
    print(json.dumps(line))
    sys.stdout.flush()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


def inference_experiment(Y_pin, w):
    """A duck-typed experiment over pinned host counts (re-uploaded by the API on every e2e step)."""
    from poisson_gpfa_b200 import util

    class _E:
        pass
    e = _E()
    Yn = Y_pin.numpy()
    e.data = [{'Y': Yn[r]} for r in range(Yn.shape[0])]
    e.Y_all = Y_pin
    e.trialDur, e.binSize = w["T"] * w["binSize"], w["binSize"]
    e.T, e.ydim, e.numTrials = w["T"], w["N"], Yn.shape[0]
    return e


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="timed steps (default: 20 for our arm, 5 for --impl reference)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--trials", type=int, default=0, help="override the trial count (debug only)")
    ap.add_argument("--cpu-trials", type=int, default=2)
    ap.add_argument("--e2e-steps", type=int, default=5)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--profile-mode", action="store_true", help="timed loop only (for runs under ncu)")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 5 if args.impl == "reference" else 20
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    args.warmup_ref = min(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
