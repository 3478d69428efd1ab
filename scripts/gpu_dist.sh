cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
# 2-GPU Model.solve variants on one box: P2P HYBRID SOLVER N SHARE_DOFS
run() {
rm -f /tmp/plf_id; for r in 0 1; do PLF_SHARE_DOFS=$5 PLF_P2P=$1 PLF_TIMING=1 PLF_HYBRID=$2 RANK=$r WORLD_SIZE=2 LOCAL_RANK=$r timeout 240 python tests/dist_model_ranks.py $4 /tmp/plf_id $3 > gpurun_out/dist_rank$r.log 2>&1 & done; wait
echo "p2p=$1 hybrid=$2 solver=$3 N=$4 share_dofs=$5"; grep "plf timing" gpurun_out/dist_rank0.log | tail -1; tail -1 gpurun_out/dist_rank0.log | cut -c1-420
tail -2 gpurun_out/dist_rank1.log | cut -c1-300
}
for N in ${DN:-1024}; do
run 1 2 replicated $N 1000000
run 1 2 replicated $N 100000
run 1 2 replicated $N 0
run 0 1 replicated $N 0
done
