#!/bin/bash
# developer sweep over NN kernel shapes (library built with -DISR_NN_TUNING)
for v in "$@"; do
  echo "== variant $v"
  ISR_NN_VARIANT=$v python scripts/perf_probe.py nn 2>&1 | grep "nn B=128"
done
