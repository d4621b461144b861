#!/bin/bash
# Round-2 measurement bundle (one gpurun call): bench lines, launch list, ncu --set full captures.
# Every ncu run follows a plain run of the same command that exited 0.  Outputs: gpurun_out/r02b_*.
set -x
O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on -f"
exp() { ncu -i $O/$1.ncu-rep --page raw --csv > $O/$1_raw.csv; ncu -i $O/$1.ncu-rep --page source --print-source cuda,sass --csv > $O/$1_src.csv 2>/dev/null; rm -f $O/$1.ncu-rep; }
timeout 900 python bench.py > $O/r02b_bench.json 2> $O/r02b_bench.err; echo bench rc $?
timeout 600 python bench.py --metric icp --skip-extra --skip-cpu --icp-iters 50 > $O/r02b_bench_icp.json 2> $O/r02b_bench_icp.err; echo bench icp rc $?
timeout 600 python bench.py --impl reference > $O/r02b_bench_reference_arm.json 2> $O/r02b_bench_reference_arm.err; echo ref rc $?
SMALL="bench.py --steps 2 --warmup 1 --candidates 64 --skip-cpu --icp-iters 2 --skip-extra"
timeout 300 python $SMALL > $O/plain_small.log 2>&1 && timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02b_launches.csv python $SMALL > $O/ncu_small.log 2>&1; echo launchlist rc $?
timeout 120 python scripts/ncu_prune.py > $O/plain_prune.log 2>&1 && timeout 600 $NCU -k regex:nn2_pruned -c 1 -o $O/r02b_nn2_pruned python scripts/ncu_prune.py > $O/ncu_prune.log 2>&1; echo rc $?; exp r02b_nn2_pruned
for w in 1 8; do timeout 120 python scripts/ncu_icp.py $w 7 > $O/plain_icp$w.log 2>&1 && timeout 600 $NCU -k regex:nn2_pruned -s 6 -c 1 -o $O/r02b_icp_fused_w$w python scripts/ncu_icp.py $w 7 > $O/ncu_icp$w.log 2>&1; echo rc $?; exp r02b_icp_fused_w$w; done
timeout 200 python scripts/ncu_hbm_kernels.py > $O/plain_hbm.log 2>&1 && timeout 900 $NCU -k regex:"transform_aos|prepare_soa7|tile_spheres_kernel|icp_accumulate_rows|transform_f64|adds_bounds" -c 14 -o $O/r02b_hbm python scripts/ncu_hbm_kernels.py > $O/ncu_hbm.log 2>&1; echo rc $?
ncu -i $O/r02b_hbm.ncu-rep --page raw --csv > $O/r02b_hbm_raw.csv; rm -f $O/r02b_hbm.ncu-rep
ls -la $O | tail -20
