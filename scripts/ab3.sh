cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 300 python scripts/check_variants.py 2>&1 | tail -18
run() {
  tag=$1; shift
  timeout 120 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --compact 1 "$@" > gpurun_out/ab3_$tag.json 2> gpurun_out/ab3_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/ab3_$tag.json'));print('$tag','resp_ms',round(d['roofline']['ms'],4),'frac',round(d['roofline']['frac'],3))" || tail -3 gpurun_out/ab3_$tag.err
}
run A_staged --opt k1_only=1 --opt passa_direct=0
run F_fm3 --opt passa_direct=0 --opt max_fm=3
run F_fm4_b3 --opt passa_direct=0 --opt max_fm=4 --opt minb=3
run F_fm4_b4 --opt passa_direct=0 --opt max_fm=4 --opt minb=4
run F_fm4_b5 --opt passa_direct=0 --opt max_fm=4 --opt minb=5
run F_fm4_b6 --opt passa_direct=0 --opt max_fm=4 --opt minb=6
run F_fm4_b4_direct --opt passa_direct=1 --opt max_fm=4 --opt minb=4
