cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "K_" 2>&1 | tail -3
run() {
  tag=$1; shift
  timeout 120 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu "$@" > gpurun_out/ab5_$tag.json 2> gpurun_out/ab5_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/ab5_$tag.json'));print('$tag','ms',round(d['ms_per_step'],4),'resp_ms',round(d['roofline']['ms'],4),'asm_ms',round(d['roofline']['assembly']['ms'],4),'asm frac',round(d['roofline']['assembly']['frac'],3))" || tail -3 gpurun_out/ab5_$tag.err
}
run asm0 --assembly-tiled 0
run asm2 --assembly-tiled 2
run asm3 --assembly-tiled 3
