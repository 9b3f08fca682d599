#!/bin/bash
# per-launch device times of one warm E-step at the headline shape (TAUSCALE sets the timescales, hence the rank)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \
    --log-file gpurun_out/launches_estep.csv python tools/prof_lowrank.py > gpurun_out/launches_estep.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open('gpurun_out/launches_estep.csv')))
hdr = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
H = rows[hdr]
kn, mn, mv, idc = H.index('Kernel Name'), H.index('Metric Name'), H.index('Metric Value'), H.index('ID')
d = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) > mv:
        d.setdefault(int(r[idc]), {'name': r[kn].split('(')[0].replace('void ', '').replace('<unnamed>::', '')})[r[mn]] = float(r[mv].replace(',', ''))
L = list(d.values())
# second E-step = after the second pcg_init... take the last half by locating lr_bins
idx = [i for i, x in enumerate(L) if x['name'].startswith('lr_bins')]
seg = L[idx[-1]:]
for x in seg:
    print('%-40s %9.1f us  %8.1f MB' % (x['name'][:40], x['gpu__time_duration.sum'] / 1e3, (x.get('dram__bytes_read.sum', 0) + x.get('dram__bytes_write.sum', 0)) / 1e6))
PY
