cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
for n in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29700+n)) bench.py --gpus $n --steps 100 --warmup 5 --no-cpu 2>/dev/null | grep '^{' > gpurun_out/scale_n$n.json
done
python bench.py --steps 100 --warmup 5 --no-cpu 2>/dev/null | grep '^{' > gpurun_out/scale_n1.json
for n in 1 2 4 8; do python -c "
import json,sys; d=json.load(open('gpurun_out/scale_n$n.json')); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"; done
