#!/bin/bash
# Round-2 closing pass (run under gpurun, one GPU): full GPU test suite (slow full-size config-5 parity included), the bench
# line and the reference arm, the launch list of the bench command, one `ncu --set full` capture of the bench workload
# (roofline.traffic), compute-sanitizer on the small configurations.
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -4 > gpurun_out/r2_gputests.log; cat gpurun_out/r2_gputests.log
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; tail -c 300 gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 0 > gpurun_out/r2_bench_reference_arm.json 2>> gpurun_out/r2_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r2_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-stress --no-batch64 > gpurun_out/bench_under_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:lm_kernel -c 1 -o /tmp/prof_c3 python tools/prof1.py 3 1 20 > gpurun_out/r2_c3_prof.log 2>&1
ncu -i /tmp/prof_c3.ncu-rep --page raw --csv > gpurun_out/r2_c3_ncu_raw.csv 2>/dev/null
ncu -i /tmp/prof_c3.ncu-rep --page details --csv > gpurun_out/r2_c3_ncu_details.csv 2>/dev/null
ncu -i /tmp/prof_c3.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > gpurun_out/r2_c3_ncu_source.csv.gz
compute-sanitizer --tool memcheck python tools/san_small.py 2>&1 | grep -v "^=========     " | tail -12 > gpurun_out/r2_sanitizer_memcheck.log
compute-sanitizer --tool racecheck python tools/san_small.py 2>&1 | grep -v "^=========     " | tail -12 > gpurun_out/r2_sanitizer_racecheck.log
cat gpurun_out/r2_sanitizer_memcheck.log gpurun_out/r2_sanitizer_racecheck.log
ls -la gpurun_out | tail -20
