cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python bench.py --steps 50 --warmup 5 --no-cpu > gpurun_out/ab8_bench.json 2> gpurun_out/ab8_bench.err
python -c "import json;d=json.load(open('gpurun_out/ab8_bench.json'));print('ms',round(d['ms_per_step'],4),'frac',round(d['roofline']['frac'],3),'e2e',d['e2e'])" || tail -5 gpurun_out/ab8_bench.err
