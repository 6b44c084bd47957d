#!/bin/bash
# ncu --set full of the table-driven verify kernel as shipped (one cold 1 M-credential call), summaries only come back
out=gpurun_out/r2y; mkdir -p $out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:"^k_ed_verify_cached" -c 2 -o $out/prof_v -f python tools/ncu_cached.py > $out/ncu_v.log 2>&1
tail -2 $out/ncu_v.log
ncu -i $out/prof_v.ncu-rep --page raw --csv > $out/prof_v_raw.csv 2>/dev/null
ncu -i $out/prof_v.ncu-rep --page source --csv --kernel-name regex:k_ed_verify_cached --print-source sass > /tmp/v_source.csv 2>/dev/null
python tools/ncu_stalls.py /tmp/v_source.csv > $out/verify_cached_stalls.txt 2>&1
python tools/ncu_summary.py $out/prof_v_raw.csv > $out/verify_cached_summary.txt 2>&1
rm -f $out/prof_v.ncu-rep
cat $out/verify_cached_summary.txt $out/verify_cached_stalls.txt
