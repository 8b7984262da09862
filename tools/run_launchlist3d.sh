#!/usr/bin/env bash
out=${1:-gpurun_out/r3a}
mkdir -p "$out"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file "$out/launches_c5_256.csv" \
  python tools/profile_kernels.py 0 256 > "$out/run.log" 2>&1; echo "rc=$?"
python - "$out/launches_c5_256.csv" <<'PY'
import csv, collections, re, sys
rows = list(csv.reader(open(sys.argv[1], errors='ignore')))
hi = [i for i, r in enumerate(rows) if 'Kernel Name' in r][0]
h = rows[hi]
kn, mv, mn, gs = h.index('Kernel Name'), h.index('Metric Value'), h.index('Metric Name'), h.index('Grid Size')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv or r[mn] != 'gpu__time_duration.sum': continue
    name = re.sub(r'\(.*', '', r[kn]).replace('void unnamed>::', '')
    v = float(r[mv].replace(',', ''))
    agg[(name, r[gs])][0] += 1; agg[(name, r[gs])][1] += v
tot = sum(v[1] for v in agg.values())
print('total ms', tot / 1e6)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:26]:
    print('%4d x %8.1f us avg  total %8.1f us  %s grid %s' % (v[0], v[1] / 1e3 / v[0], v[1] / 1e3, k[0][:44], k[1]))
PY
