set -x
timeout 600 python -m pytest tests -q -m gpu 2>&1 | tail -3
timeout 300 python tools/prof1.py 5 2 3 1 2>/dev/null | grep "cfg\|p\."
timeout 300 python tools/prof1.py 5 2 3 0 2>/dev/null | grep "cfg"
ncu --set full --import-source on --clock-control none -k regex:lm_kernel -c 1 -o /tmp/prof_c5 python tools/prof1.py 5 1 2 > gpurun_out/prof_r1_c5.log 2>&1
ncu -i /tmp/prof_c5.ncu-rep --page source --csv --print-source cuda,sass 2>/dev/null | gzip > gpurun_out/r1_c5_ncu_source.csv.gz
ls -la gpurun_out | tail -5
