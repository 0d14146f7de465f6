#!/bin/bash
# Round-1 profiling pass (run under gpurun): tests, launch list of the bench command, full ncu captures of the
# persistent kernel on config 3 (latency-bound, L2-resident) and config 5 (HBM-bound stress graph).
# The .ncu-rep files are exported to CSV on the box (raw / details / source pages) so the copy-back stays small.
set -x
mkdir -p gpurun_out

ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-stress --no-batch64 > gpurun_out/bench_under_ncu.log 2>&1
for cfg in 3 5; do
  if [ $cfg = 3 ]; then args="3 1 20"; else args="5 1 2"; fi
  ncu --set full --import-source on --clock-control none -k regex:lm_kernel -c 1 -o /tmp/prof_c$cfg python tools/prof1.py $args > gpurun_out/prof_r2_c$cfg.log 2>&1
  ncu -i /tmp/prof_c$cfg.ncu-rep --page raw --csv > gpurun_out/r2_c${cfg}_ncu_raw.csv 2>/dev/null
  ncu -i /tmp/prof_c$cfg.ncu-rep --page details --csv > gpurun_out/r2_c${cfg}_ncu_details.csv 2>/dev/null
  ncu -i /tmp/prof_c$cfg.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > gpurun_out/r2_c${cfg}_ncu_source.csv.gz
done
python tools/prof1.py 5 2 3 > gpurun_out/c5_timing.log 2>&1
tail -25 gpurun_out/c5_timing.log
ls -la gpurun_out
