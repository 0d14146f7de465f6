#!/bin/bash
# quick GPU check used while tuning the PCG loops: gpu tests, then timed solves of configs 3 / 2 / 5 (resident and old loop)
timeout 400 python -m pytest tests -m gpu -q -x > gpurun_out/t.log 2>&1; tail -3 gpurun_out/t.log
python tools/prof1.py 3 2 20 1 > gpurun_out/p3.log 2>&1; python tools/prof1.py 2 2 20 1 > gpurun_out/p2.log 2>&1
python tools/prof1.py 3 2 20 0 > gpurun_out/p3n.log 2>&1;  python tools/prof1.py 3 2 20 32 > gpurun_out/p3old.log 2>&1
python tools/prof1.py 2 2 20 0 > gpurun_out/p2n.log 2>&1; python tools/prof1.py 2 2 20 32 > gpurun_out/p2old.log 2>&1; python tools/prof1.py 5 2 3 0 > gpurun_out/p5n.log 2>&1
cat gpurun_out/p3.log; tail -12 gpurun_out/p2.log; head -1 gpurun_out/p3n.log gpurun_out/p3old.log gpurun_out/p2n.log gpurun_out/p2old.log gpurun_out/p5n.log
python tools/prof1.py 3 2 20 64 2>&1 | grep "linearize\|probe" ; python tools/prof1.py 2 2 20 64 2>&1 | grep "linearize\|probe"
