cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 60 scripts/bin/dfma_lat > gpurun_out/dfma_lat.txt 2>&1; cat gpurun_out/dfma_lat.txt
timeout 300 python scripts/check_variants.py 2>&1 | tail -8
run() {
  tag=$1; shift
  timeout 120 python bench.py --steps 100 --warmup 5 --no-e2e --no-cpu --compact 1 "$@" > gpurun_out/ab2_$tag.json 2> gpurun_out/ab2_$tag.err
  python -c "import json;d=json.load(open('gpurun_out/ab2_$tag.json'));print('$tag','resp_ms',round(d['roofline']['ms'],4))" || tail -3 gpurun_out/ab2_$tag.err
}
run A_staged --opt k1_only=1 --opt passa_direct=0
run A_direct --opt k1_only=1 --opt passa_direct=1
run A_direct_b4 --opt k1_only=1 --opt passa_direct=1 --opt minb=4
run A_warp2_b3 --opt k1_only=1 --opt passa_warp=2 --opt minb_a=3
run A_warp3_b3 --opt k1_only=1 --opt passa_warp=3 --opt minb_a=3
run A_warp2_b4 --opt k1_only=1 --opt passa_warp=2 --opt minb_a=4
run A_warp3_b4 --opt k1_only=1 --opt passa_warp=3 --opt minb_a=4
run A_warp4_b4 --opt k1_only=1 --opt passa_warp=4 --opt minb_a=4
run A_warp3_b5 --opt k1_only=1 --opt passa_warp=3 --opt minb_a=5
run F_staged_B3 --opt passa_direct=0 --opt minb=3
run F_staged_B4 --opt passa_direct=0 --opt minb=4
run F_warp3b4_B3 --opt passa_warp=3 --opt minb_a=4 --opt minb=3
run F_warp3b4_B4 --opt passa_warp=3 --opt minb_a=4 --opt minb=4
