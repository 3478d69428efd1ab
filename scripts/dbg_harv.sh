cd $GRAFT_REPO_ROOT
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
python bench.py --steps 100 --warmup 5 --no-cpu > gpurun_out/harv_bench.json 2> gpurun_out/harv_bench.err
python -c "
import json; d=json.loads(open('gpurun_out/harv_bench.json').read()); print(d['value'], d['ms_per_step']); print(json.dumps(d.get('harvested_mix'))[:3000]); print(json.dumps(d['e2e_load_step'])[:300])"; tail -3 gpurun_out/harv_bench.err
