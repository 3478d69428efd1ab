# A/B of the return-mapping kernel variants + parity tests
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/fused_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/fused_tests.log
tail -5 gpurun_out/fused_tests.log
run() {
  tag=$1; shift
  timeout 300 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu "$@" > gpurun_out/ab_$tag.json 2> gpurun_out/ab_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/ab_$tag.json'));print('$tag','ms',round(d['ms_per_step'],4),'resp_ms',round(d['roofline']['ms'],4),'frac',round(d['roofline']['frac'],3),'asm_ms',round(d['roofline']['assembly']['ms'],4))" || tail -3 gpurun_out/ab_$tag.err
}
run c1_staged_fm2 --compact 1 --opt passa_direct=0 --opt max_fm=2
run c1_direct_fm2 --compact 1 --opt passa_direct=1 --opt max_fm=2
run c1_direct_fm3 --compact 1 --opt passa_direct=1 --opt max_fm=3
run c1_direct_fm3_b4 --compact 1 --opt passa_direct=1 --opt max_fm=3 --opt minb=4
run c1_staged_fm3 --compact 1 --opt passa_direct=0 --opt max_fm=3
run c2_direct_fm3 --compact 2 --opt passa_direct=1 --opt max_fm=3
run c2_direct_fm3_b4 --compact 2 --opt passa_direct=1 --opt max_fm=3 --opt minb=4
run c2_staged_fm2 --compact 2 --opt passa_direct=0 --opt max_fm=2
run c2_direct_fm2 --compact 2 --opt passa_direct=1 --opt max_fm=2
