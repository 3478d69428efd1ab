cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "fused or multigrid or inclusion" 2>&1 | tail -6
for fuse in 2 1 0; do
PLF_MG_FUSE_TAIL=$fuse timeout 600 python scripts/solve_demo.py 256 5 0.004 2>&1 | grep -E "^N=|linear solves|launches"
PLF_MG_FUSE_TAIL=$fuse timeout 600 python scripts/solve_demo.py 1024 5 0.004 2>&1 | grep -E "^N=|linear solves|launches"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 20000 -c 600 --csv --log-file gpurun_out/r2_solve1024_launches.csv python scripts/solve_demo.py 1024 5 0.004 > gpurun_out/r2_solve1024.log 2>&1
tail -2 gpurun_out/r2_solve1024.log
