#!/bin/bash
# ncu --set full of the four-lane verify kernel (mode given as $1: 1 = after k_ed_hram, 2 = fused), summaries only come back
m=${1:-1}; out=gpurun_out/r2q; mkdir -p $out
AFC_VERIFY_QUAD=$m timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"^k_ed_(verify_quad|quad_finish)" -c 4 -o $out/prof_quad$m -f python tools/ncu_cached.py > $out/ncu_quad$m.log 2>&1
tail -2 $out/ncu_quad$m.log
ncu -i $out/prof_quad$m.ncu-rep --page raw --csv > $out/prof_quad${m}_raw.csv 2>/dev/null
ncu -i $out/prof_quad$m.ncu-rep --page source --csv --kernel-name regex:k_ed_verify_quad --print-source sass > /tmp/quad_source.csv 2>/dev/null
python tools/ncu_stalls.py /tmp/quad_source.csv > $out/quad${m}_stalls.txt 2>&1
python tools/ncu_summary.py $out/prof_quad${m}_raw.csv > $out/quad${m}_summary.txt 2>&1
rm -f $out/prof_quad$m.ncu-rep
ls -la $out | tail -8
