#!/bin/bash
# Final build of round 2 (packed resolve pass 1): one ncu --set full capture of the verification search,
# after a plain run of the same program that exited 0.  Outputs: gpurun_out/r02d_*.
O=gpurun_out
timeout 40 python scripts/ncu_prune.py > $O/plain_prune.log 2>&1 || exit 1
timeout 60 ncu --set full --clock-control none --import-source on -f -k regex:nn2_pruned -c 1 -o $O/r02d_nn2_pruned python scripts/ncu_prune.py > $O/ncu_prune.log 2>&1
ncu -i $O/r02d_nn2_pruned.ncu-rep --page raw --csv > $O/r02d_nn2_pruned_raw.csv
ncu -i $O/r02d_nn2_pruned.ncu-rep --page source --print-source cuda,sass --csv > $O/r02d_nn2_pruned_src.csv 2>/dev/null
rm -f $O/r02d_nn2_pruned.ncu-rep
