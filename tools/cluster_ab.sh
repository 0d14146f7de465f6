#!/bin/bash
# A/B of the team modes on the small graphs: global-counter barrier (PUS_CLUSTER=0) vs one cluster per team (8 / 16)
for m in 0 8 16; do
  echo "== PUS_CLUSTER=$m"
  PUS_CLUSTER=$m timeout 120 python tools/prof1.py 2 3 20 0 | head -1
  PUS_CLUSTER=$m timeout 120 python tools/batch_quick.py 8
  PUS_CLUSTER=$m timeout 120 python tools/batch_quick.py 16
  PUS_CLUSTER=$m timeout 120 python tools/batch_quick.py 64
done
