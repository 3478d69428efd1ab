cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus $NG"
# 1024^2 J2 load sequence on all GPUs, level 0 + CG shared, exchanges through peer mailboxes; rank 0 compares with one GPU
rm -f /tmp/plf_id; for r in $(seq 0 $((NG-1))); do PLF_TIMING=1 RANK=$r WORLD_SIZE=$NG LOCAL_RANK=$r timeout 240 python tests/dist_model_ranks.py ${DN:-1024} /tmp/plf_id replicated > gpurun_out/hyb_rank$r.log 2>&1 & done; wait
tail -1 gpurun_out/hyb_rank0.log | tee gpurun_out/r2_hybrid_${NG}gpu_1024.json | cut -c1-700; tail -1 gpurun_out/hyb_rank$((NG-1)).log | cut -c1-200
# BASELINE config 5 the same way
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29553 scripts/config5_run.py 2048 2 0.001 2> gpurun_out/r2_config5_${NG}gpu_hybrid.err | grep '^{' > gpurun_out/r2_config5_${NG}gpu_hybrid.json
cat gpurun_out/r2_config5_${NG}gpu_hybrid.json | cut -c1-1200; tail -2 gpurun_out/r2_config5_${NG}gpu_hybrid.err | cut -c1-300
