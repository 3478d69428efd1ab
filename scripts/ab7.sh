cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
run() {
  tag=$1; shift
  timeout 120 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu "$@" > gpurun_out/ab7_$tag.json 2> gpurun_out/ab7_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/ab7_$tag.json'));print('$tag','resp_ms',round(d['roofline']['ms'],4))" || tail -3 gpurun_out/ab7_$tag.err
}
run A_only --opt k1_only=1
run A_nomath_staged --opt k1_only=2
run A_nomath_direct --opt k1_only=2 --opt passa_direct=1
