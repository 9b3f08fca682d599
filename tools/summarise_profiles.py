"""Turn the raw ncu captures of tools/profile_round.sh into the tracked summaries under profiles/.

  python tools/summarise_profiles.py [--tag r01] [--src gpurun_out]

* <src>/launches_bench.csv (ncu launch list of `bench.py --steps 1 --warmup 3 --profile-mode`)
    -> profiles/<tag>_launches_bench.csv        (copy)
    -> profiles/<tag>_launches_bench_step.json  (per-kernel time / DRAM bytes of the last EM iteration)
    -> profiles/<tag>_factor_traffic.json       (DRAM bytes of one batched factorisation; read by bench.py)
* <src>/prof_panel.ncu-rep (ncu --set full of two mid-factorisation chol_panel_kernel launches)
    -> profiles/<tag>_ncu_panel_current.txt
Per-launch times in the launch list are serialised and cold-cache: compare shares, not absolutes.
"""
import argparse
import collections
import csv
import json
import os
import shutil
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

PANEL_METRICS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed.sum", "sm__cycles_elapsed.max",
]


def short(name):
    return name.split('(')[0].replace('void ', '').replace('<unnamed>::', '')


def read_launches(path):
    rows = list(csv.reader(open(path)))
    hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    H = rows[hdr]
    kn, mn, mv, idc = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('ID')
    d = collections.OrderedDict()
    for r in rows[hdr + 1:]:
        if len(r) > mv:
            d.setdefault(int(r[idc]), {'name': short(r[kn])})[r[mn]] = float(r[mv].replace(',', ''))
    return list(d.values())


def launch_summaries(src, tag):
    path = os.path.join(src, "launches_bench.csv")
    L = read_launches(path)
    shutil.copy(path, os.path.join(ROOT, "profiles", tag + "_launches_bench.csv"))
    ms = lambda x: x['gpu__time_duration.sum'] / 1e6
    by = lambda x: x.get('dram__bytes_read.sum', 0.0) + x.get('dram__bytes_write.sum', 0.0)
    piv = [i for i, x in enumerate(L) if x['name'].startswith('pivchol')]
    lowrank = len(piv) >= 2
    if lowrank:     # low-rank posterior pass: one EM iteration = from one prior factorisation (pivchol) to the next
        step = L[piv[-2]:piv[-1]]
    else:           # dense: from after the selected inverse of iteration k-1 to the selected inverse of iteration k
        big = [i for i, x in enumerate(L) if x['name'].startswith('lauum_tiles') and ms(x) > 5.0]
        step = L[big[-2] + 1:big[-1] + 1]
    agg = collections.OrderedDict()
    for x in step:
        a = agg.setdefault(x['name'], [0, 0.0, 0.0])
        a[0] += 1; a[1] += ms(x); a[2] += by(x)
    tot = sum(a[1] for a in agg.values())
    out = {"what": "ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none of "
                   "`python bench.py --steps 1 --warmup 3 --profile-mode`; one steady-state EM iteration (M-step of "
                   "iteration k-1 + E-step of iteration k); per-launch times are serialised/cold-cache: compare shares",
           "launches_in_step": len(step), "sum_ms": tot,
           "kernels": [{"kernel": k, "launches": v[0], "ms": round(v[1], 3), "share": round(v[1] / tot, 4), "dram_bytes": v[2]}
                       for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
    json.dump(out, open(os.path.join(ROOT, "profiles", tag + "_launches_bench_step.json"), "w"), indent=1)
    print("step: %d launches, %.1f ms serialised" % (len(step), tot))
    for k in out["kernels"][:10]:
        print("  %-36s x%-4d %8.2f ms  %5.1f%%" % (k["kernel"], k["launches"], k["ms"], 100 * k["share"]))
    if lowrank:
        return
    # the factorisation at the mode: the chol_diag + chol_panel launches right before the polish solve
    isolve = max(i for i, x in enumerate(L) if x['name'].startswith('chol_solve_kernel<double>'))
    j = isolve - 1
    fac = []
    while j >= 0 and (L[j]['name'].startswith('chol_panel') or L[j]['name'].startswith('chol_diag')):
        fac.append(L[j]); j -= 1
    traffic = {"what": "DRAM traffic of ONE batched factorisation call of the bench workload (chol_diag_kernel + chol_panel_kernel "
                       "launches, all stream parts) from the ncu launch list of the bench command (profiles/%s_launches_bench.csv)" % tag,
               "dram_bytes_per_factorisation": sum(by(x) for x in fac), "serialised_ms": sum(ms(x) for x in fac),
               "launches": len(fac),
               "note": "operands: each panel CTA streams 2*j tiles of 32 KB; the L(j,:) row panel is shared by a slot's CTAs "
                       "through L2; outputs L (FP64) + FP32 mirror"}
    json.dump(traffic, open(os.path.join(ROOT, "profiles", tag + "_factor_traffic.json"), "w"), indent=1)
    print("factorisation: %d launches, %.1f GB" % (len(fac), traffic["dram_bytes_per_factorisation"] / 1e9))


def gemm_summary(src, tag):
    """ncu --set full of the three gemm_nt launches of a warm E-step (capacitance, Yh, post_vsmGP)."""
    rep = os.path.join(src, "prof_gemm_nt.ncu-rep")
    if not os.path.exists(rep):
        print("no", rep)
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units = rows[0], rows[1]
    extra = ["smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
             "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]
    lines = ["# ncu --set full --clock-control none --import-source on -k regex:gemm_nt, tools/prof_lowrank.py (1024 trials, "
             "q=8, T=200): capacitance blocks, Yh, post_vsmGP launches of a warm E-step"]
    last = None
    for r in rows[2:]:
        lines.append("--- kernel %s grid %s" % (r[H.index("Kernel Name")][:40], r[H.index("Grid Size")]))
        for m in PANEL_METRICS + extra:
            if m in H:
                lines.append("   %-80s %s %s" % (m, r[H.index(m)], units[H.index(m)]))
        last = r
    open(os.path.join(ROOT, "profiles", tag + "_ncu_gemm_nt.txt"), "w").write("\n".join(lines) + "\n")
    def val(m, scale=1.0):
        u = units[H.index(m)]
        v = float(last[H.index(m)].replace(",", ""))
        return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0) * scale
    rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
    json.dump({"what": "DRAM traffic of the post_vsmGP launch of gemm_nt_kernel (q x trials symmetric T x T x r products) from the "
                       "ncu --set full capture of tools/prof_lowrank.py (profiles/%s_ncu_gemm_nt.txt, last launch), 1024 trials" % tag,
               "dram_bytes_per_launch": rd + wr, "dram_read": rd, "dram_write": wr,
               "tensor_pipe_active_pct": float(last[H.index("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")]),
               "grid": last[H.index("Grid Size")]},
              open(os.path.join(ROOT, "profiles", tag + "_syrk_traffic.json"), "w"), indent=1)
    print("\n".join(lines[-24:]))


def panel_summary(src, tag, header):
    rep = os.path.join(src, "prof_panel.ncu-rep")
    if not os.path.exists(rep):
        print("no", rep)
        return
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units = rows[0], rows[1]
    lines = ["# " + header]
    for r in rows[2:]:
        lines.append("--- kernel " + r[H.index("Kernel Name")])
        for m in PANEL_METRICS:
            if m in H:
                lines.append("   %-72s %s %s" % (m, r[H.index(m)], units[H.index(m)]))
    open(os.path.join(ROOT, "profiles", tag + "_ncu_panel_current.txt"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[:12]))


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--tag", default="r01")
    ap.add_argument("--src", default=os.path.join(ROOT, "gpurun_out"))
    ap.add_argument("--header", default="ncu --set full --clock-control none --import-source on, tools/prof_factor.py "
                                        "(1024 trials, q=8, T=200, n=1600), current build")
    a = ap.parse_args()
    launch_summaries(a.src, a.tag)
    gemm_summary(a.src, a.tag)
    if os.environ.get("DENSE"):
        panel_summary(a.src, a.tag, a.header)
