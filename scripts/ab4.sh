cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 60 scripts/bin/membw > gpurun_out/membw.txt 2>&1; cat gpurun_out/membw.txt
timeout 300 python scripts/check_variants.py 2>&1 | tail -14
run() {
  tag=$1; shift
  timeout 120 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --compact 1 "$@" > gpurun_out/ab4_$tag.json 2> gpurun_out/ab4_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/ab4_$tag.json'));print('$tag','resp_ms',round(d['roofline']['ms'],4),'frac',round(d['roofline']['frac'],3))" || tail -3 gpurun_out/ab4_$tag.err
}
run A_staged --opt k1_only=1 --opt passa_direct=0
run F_fm4_staged --opt passa_direct=0
run F_fm4_direct --opt passa_direct=1
run F_fm3_staged --opt passa_direct=0 --opt max_fm=3
