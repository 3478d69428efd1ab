cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
NCCL_DEBUG=INFO timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 40 --warmup 3 --no-cpu > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
grep -E "NCCL INFO.*(via|P2P|SHM|NVLS|Channel 00)" gpurun_out/r2_bench_2gpu.err | head -8
tail -3 gpurun_out/r2_bench_2gpu.err | cut -c1-300
python -c "
import json; d=json.loads([l for l in open('gpurun_out/r2_bench_2gpu.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'])
print(json.dumps(d.get('coupled_solve'))[:1000])
"
