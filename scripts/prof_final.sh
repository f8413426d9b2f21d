#!/bin/bash
# Final-build evidence (run under gpurun, one GPU): (1) the ncu launch list of the default bench command's timed region,
# (2) one `--set full` capture of msm_accumulate_kernel (96 uniform columns x 2^16) and of ntt_pass_kernel (64 columns x 2^16,
# lagrange_to_coeff: both passes).  Every program is run once WITHOUT ncu first and must exit 0.
set -u
out=gpurun_out
tag=${1:-r02_final}
python bench.py --steps 2 --warmup 1 --no-extras > $out/${tag}_plain_bench.json 2> $out/${tag}_plain_bench.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-extras > /dev/null 2>&1
python - <<PY
import csv, collections
rows = [r for r in csv.reader(open('$out/${tag}_launches.csv')) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    agg.setdefault(r[ki].split('(')[0], []).append(v)
tot = sum(sum(v) for v in agg.values())
with open('$out/${tag}_launches.txt', 'w') as f:
    f.write('# ncu --metrics gpu__time_duration.sum --clock-control none: python bench.py --steps 2 --warmup 1 --no-extras (cold-cache, serialised launches)\n')
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        f.write('%-50s n=%4d  mean=%10.1f us  sum=%10.1f us  share=%5.1f %%\n' % (k[:50], len(v), sum(v) / len(v) / 1e3, sum(v) / 1e3, 100 * sum(v) / tot))
print(open('$out/${tag}_launches.txt').read())
PY
COLS=96 DIST=uniform python scripts/prof_msm.py > $out/${tag}_plain_msm.log 2>&1 || exit 1
COLS=96 DIST=uniform ncu --set full --clock-control none --import-source on -k regex:msm_accumulate -c 1 -f -o $out/${tag}_acc \
    python scripts/prof_msm.py > /dev/null 2>&1
ncu -i $out/${tag}_acc.ncu-rep --page details > $out/${tag}_acc_details.txt 2>&1
ncu -i $out/${tag}_acc.ncu-rep --page raw --csv > $out/${tag}_acc_raw.csv 2>&1
COLS=64 python scripts/prof_ntt.py > $out/${tag}_plain_ntt.log 2>&1 || exit 1
COLS=64 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -c 2 -f -o $out/${tag}_ntt \
    python scripts/prof_ntt.py > /dev/null 2>&1
ncu -i $out/${tag}_ntt.ncu-rep --page details > $out/${tag}_ntt_details.txt 2>&1
ncu -i $out/${tag}_ntt.ncu-rep --page raw --csv > $out/${tag}_ntt_raw.csv 2>&1
# the .ncu-rep files (≈ 40 MB each) would push gpurun_out/ over its 64 MiB return limit: the text / csv pages are what is kept
ncu -i $out/${tag}_acc.ncu-rep --page source --csv > $out/${tag}_acc_source.csv 2>&1
ncu -i $out/${tag}_ntt.ncu-rep --page source --csv > $out/${tag}_ntt_source.csv 2>&1
rm -f $out/${tag}_acc.ncu-rep $out/${tag}_ntt.ncu-rep
ls -la $out/${tag}_*
