cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
NG=$(nvidia-smi -L | wc -l); echo "gpus $NG"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $NG --steps 40 --warmup 3 --no-cpu 2> gpurun_out/r2_bench_${NG}gpu.err | grep '^{' > gpurun_out/r2_bench_${NG}gpu.json
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_${NG}gpu.json').read())
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
print(json.dumps(d.get('coupled_solve'))[:900])
"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $NG --master-addr 127.0.0.1 --master-port 29553 scripts/config5_run.py 2048 2 0.001 2> gpurun_out/r2_config5_${NG}gpu.err | grep '^{' > gpurun_out/r2_config5_${NG}gpu.json
cat gpurun_out/r2_config5_${NG}gpu.json | cut -c1-1200; tail -2 gpurun_out/r2_config5_${NG}gpu.err | cut -c1-300
