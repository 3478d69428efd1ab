cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 500 python -m pytest tests/test_gpu_model.py -q -x -k "two_gpus or partitioned" 2>&1 | tail -3
bash scripts/gpu_scale2.sh 2>&1 | grep -v "NCCL INFO"
