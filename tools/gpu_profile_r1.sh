#!/bin/bash
# Round-1 profiling pass (run under gpurun): tests, launch list of the bench command, full ncu captures of the
# persistent kernel on config 3 (latency-bound, L2-resident) and config 5 (HBM-bound stress graph).
set -x
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:lm_kernel -c 1 -o gpurun_out/prof_r1_c3 python tools/prof1.py 3 1 20 > gpurun_out/prof_r1_c3.log 2>&1
python tools/prof1.py 5 2 3 > gpurun_out/c5_timing.log 2>&1
tail -25 gpurun_out/c5_timing.log
ncu --set full --import-source on --clock-control none -k regex:lm_kernel -c 1 -o gpurun_out/prof_r1_c5 python tools/prof1.py 5 1 2 > gpurun_out/prof_r1_c5.log 2>&1
ls -la gpurun_out
