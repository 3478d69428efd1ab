cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python scripts/check_variants.py 2>&1 | tail -14
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
run() {
  tag=$1; shift
  timeout 120 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu "$@" > gpurun_out/ab6_$tag.json 2> gpurun_out/ab6_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/ab6_$tag.json'));print('$tag','ms',round(d['ms_per_step'],4),'resp_ms',round(d['roofline']['ms'],4),'frac',round(d['roofline']['frac'],3),'asm_ms',round(d['roofline']['assembly']['ms'],4))" || tail -3 gpurun_out/ab6_$tag.err
}
run A_only --opt k1_only=1
run full
run full_direct --opt passa_direct=1
run fm2 --opt max_fm=2
