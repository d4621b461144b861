set -x
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 500 python bench.py > gpurun_out/bench_r01_h.json 2> gpurun_out/bench_r01_h.err; echo bench rc $?
timeout 400 python bench.py --impl reference > gpurun_out/bench_ref_h.json 2> gpurun_out/bench_ref_h.err; echo ref rc $?
timeout 300 python bench.py --steps 2 --warmup 1 --candidates 64 --skip-cpu --icp-iters 2 --skip-extra > gpurun_out/plain_small.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_launches_pruned.csv python bench.py --steps 2 --warmup 1 --candidates 64 --skip-cpu --icp-iters 2 --skip-extra > gpurun_out/ncu_small.log 2>&1; echo launchlist rc $?
timeout 120 python scripts/ncu_prune.py > gpurun_out/plain_prune.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:nn2_pruned -c 1 -f -o gpurun_out/nn2p_r01_i python scripts/ncu_prune.py > gpurun_out/ncu_prune.log 2>&1; echo ncu rc $?
