cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for H in 1 0; do for N in ${DN:-1024}; do
rm -f /tmp/plf_id; for r in 0 1; do PLF_TIMING=1 PLF_HYBRID=$H RANK=$r WORLD_SIZE=2 LOCAL_RANK=$r timeout 200 python tests/dist_model_ranks.py $N /tmp/plf_id replicated > gpurun_out/dist_rank$r.log 2>&1 & done; wait
echo "hybrid=$H N=$N"; grep "plf timing" gpurun_out/dist_rank0.log | tail -1; tail -1 gpurun_out/dist_rank0.log | cut -c1-330
done; done
nvidia-smi topo -m | head -12
