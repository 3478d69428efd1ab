cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -3
PLF_MG=1 timeout 1200 python scripts/solve_demo.py 1024 5 0.004 2>&1 | grep -E "^N=|linear solves|increments|global|launches" | tee gpurun_out/solve_demo_1024_mg.txt
PLF_MG=1 timeout 1200 python scripts/solve_demo.py 256 5 0.004 2>&1 | grep -E "^N=|linear solves"
