cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py -m gpu -x -q 2>&1 | tail -5
PLF_MG=1 timeout 600 python scripts/solve_demo.py 64 5 0.004 2>&1 | grep -E "^N=|linear solves|global"
PLF_MG=0 timeout 600 python scripts/solve_demo.py 64 5 0.004 2>&1 | grep -E "^N=|linear solves|global"
PLF_MG=1 timeout 900 python scripts/solve_demo.py 256 5 0.004 2>&1 | grep -E "^N=|linear solves|global"
