# GPU checks that did not fit into round 1's GPU budget; run first thing in the next round.
#   gpurun --gpus 2 --timeout 900 -- 'bash scripts/round_start_gpu_checks.sh'
# 1. the opt-in NCCL test (whole plastic load sequence through the distributed CG with one rank; pages in libnccl: ~3 min)
# 2. the 2-rank distributed solve against the single-GPU solve (scripts/dist_solve.py)
# 3. every pass A / pass B variant against the general kernels on the full-size workload (scripts/check_variants.py)
# 4. the round profile (tests, ncu launch list, ncu --set full of the top kernels, bench, reference arm, 1024^2 demo)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
PLF_TEST_NCCL=1 timeout 400 python -m pytest tests/test_gpu_model.py -q -m gpu -k distributed_cg 2>&1 | tail -3
NG=$(nvidia-smi -L | wc -l)
if [ "$NG" -ge 2 ]; then
  timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29811 \
      scripts/dist_solve.py 512 512 > gpurun_out/dist_solve.log 2>&1; echo "dist_solve rc=$?"; grep '^{' gpurun_out/dist_solve.log | cut -c1-400
fi
timeout 200 python scripts/check_variants.py 2>&1 | tail -12
TAG=${TAG:-r2_start} timeout 600 bash scripts/prof_final.sh 2>&1 | tail -25
