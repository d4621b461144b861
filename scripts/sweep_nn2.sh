#!/bin/bash
for v in "$@"; do
  echo "== variant $v sort=$PROBE_SORT nores=$ISR_NN2_NORESOLVE"
  ISR_NN_VARIANT=$v python scripts/perf_probe.py nn 2>&1 | grep "nn exact B=128 idx=False"
done
