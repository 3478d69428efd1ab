cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_gpu_model.py -m gpu -x -q -k "two_gpus or partitioned or one_rank" 2>&1 | tail -6
for N in ${DN:-256 1024}; do
rm -f /tmp/plf_id; for r in 0 1; do RANK=$r WORLD_SIZE=2 LOCAL_RANK=$r timeout 240 python tests/dist_model_ranks.py $N /tmp/plf_id replicated > gpurun_out/dist_rank$r.log 2>&1 & done; wait
tail -1 gpurun_out/dist_rank0.log | cut -c1-900; tail -2 gpurun_out/dist_rank1.log | cut -c1-300
done
