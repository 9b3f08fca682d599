"""Per-launch times of the low-rank posterior kernels from an ncu launch list (gpurun_out/launches_bench.csv)."""
import csv, collections, sys
def short(name): return name.split('(')[0].replace('void ', '').replace('<unnamed>::', '')
rows = list(csv.reader(open(sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/launches_bench.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]; kn, mn, mv, idc = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('ID')
gs = H.index('Grid Size')
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > mv:
        e = d.setdefault(int(r[idc]), {'name': short(r[kn]), 'grid': r[gs]}); e[r[mn]] = float(r[mv].replace(',', ''))
L = list(d.values())
piv = [i for i, x in enumerate(L) if x['name'].startswith('pivchol')]
step = L[piv[-2]:piv[-1]]
agg = collections.OrderedDict()
for x in step:
    a = agg.setdefault(x['name'][:44], [0, 0.0]); a[0] += 1; a[1] += x['gpu__time_duration.sum'] / 1e6
tot = sum(a[1] for a in agg.values())
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]: print(f"{k:46s} n={v[0]:4d} {v[1]:8.3f} ms {100*v[1]/tot:5.1f}%")
print('total', round(tot, 2))
for x in step:
    if x['name'].startswith(('gemm_nt', 'lr_', 'zt_')):
        print(f"{x['name'][:30]:32s} {x['grid']:>20s} {x['gpu__time_duration.sum']/1e3:9.1f} us")
