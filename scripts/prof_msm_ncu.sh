#!/bin/bash
# per-launch device times of one short MSM workload (ncu launch list), summarised per kernel; run under gpurun
# usage: prof_msm_ncu.sh <tag>   (env DIST/K/COLS as scripts/prof_msm.py)
tag=$1
python scripts/prof_msm.py > gpurun_out/plain_$tag.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file /tmp/launch_$tag.csv python scripts/prof_msm.py > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('/tmp/launch_$tag.csv')) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    agg.setdefault(r[ki].split('(')[0], []).append(v)
with open('gpurun_out/launches_$tag.txt', 'w') as f:
    for k, v in agg.items():
        f.write('%-60s n=%3d  last=%9.1f us  mean=%9.1f us\n' % (k[:60], len(v), v[-1] / 1e3, sum(v) / len(v) / 1e3))
print(open('gpurun_out/launches_$tag.txt').read())
PY
