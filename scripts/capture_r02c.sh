#!/bin/bash
# Last build of round 2 (two-slot ring, 21 CTAs per SM for the plain search): the two-rank parity test,
# the launch list of a reduced bench run and one ncu --set full capture of the verification search.
# Every ncu run follows a plain run of the same command that exited 0.  Outputs: gpurun_out/r02c_*.
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
exp() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv; ncu -i $O/$1.ncu-rep --page source --print-source cuda,sass --csv > $O/$1_src.csv 2>/dev/null; rm -f $O/$1.ncu-rep; }
timeout 400 python -m pytest tests/test_dist_gpu.py -m gpu -x -q > $O/r02c_tests.log 2>&1; echo tests rc $?
timeout 120 python scripts/ncu_prune.py > $O/plain_prune.log 2>&1 && timeout 400 $NCU -k regex:nn2_pruned -c 1 -o $O/r02c_nn2_pruned python scripts/ncu_prune.py > $O/ncu_prune.log 2>&1; echo rc $?; exp r02c_nn2_pruned
SMALL="bench.py --steps 2 --warmup 1 --candidates 64 --skip-cpu --icp-iters 2 --skip-extra"
timeout 200 python $SMALL > $O/plain_small.log 2>&1 && timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02c_launches.csv python $SMALL > $O/ncu_small.log 2>&1; echo launchlist rc $?
