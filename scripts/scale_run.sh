# weak-scaling run of bench.py on one box (strips of one N*1024 x 1024 mesh): gpurun --gpus 8 -- 'bash scripts/scale_run.sh'
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for n in 8 4 2; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 100 --warmup 5 --no-cpu 2>gpurun_out/scale_n$n.err | grep '^{' > gpurun_out/scale_n$n.json
done
timeout 300 python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | grep '^{' > gpurun_out/scale_n1.json
for n in 1 2 4 8; do python -c "
import json,sys; d=json.load(open('gpurun_out/scale_n$n.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"; done
