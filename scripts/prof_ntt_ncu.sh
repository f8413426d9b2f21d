#!/bin/bash
# ncu --set full with source correlation of both passes of lagrange_to_coeff 2^16 x 64 (scripts/prof_ntt.py); run under gpurun.
set -u
out=gpurun_out
tag=${1:-ntt}
COLS=64 python scripts/prof_ntt.py > $out/${tag}_plain.log 2>&1 || exit 1
COLS=64 ncu --set full --clock-control none --import-source on -k regex:ntt_pass -c 2 -f -o $out/${tag} python scripts/prof_ntt.py > /dev/null 2>&1
ncu -i $out/${tag}.ncu-rep --page details > $out/${tag}_details.txt 2>&1
ncu -i $out/${tag}.ncu-rep --page raw --csv > $out/${tag}_raw.csv 2>&1
ncu -i $out/${tag}.ncu-rep --page source --csv > $out/${tag}_source.csv 2>&1
rm -f $out/${tag}.ncu-rep
ls -la $out/${tag}_*
