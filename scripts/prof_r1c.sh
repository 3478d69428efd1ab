cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'assemble' -s 3 -c 1 -o gpurun_out/r1c_asm_tiled python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --assembly-tiled 1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'assemble' -s 3 -c 1 -o gpurun_out/r1c_asm_rows python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --assembly-tiled 0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'response_pass' -s 6 -c 2 -o gpurun_out/r1c_resp python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
for t in 0 1; do python bench.py --steps 50 --warmup 5 --no-cpu --no-e2e --assembly-tiled $t 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('step ms %.4f  resp ms %.4f (frac %.3f)  asm ms %.4f (frac %.3f)'%(d['ms_per_step'],d['roofline']['ms'],d['roofline']['frac'],d['roofline']['assembly']['ms'],d['roofline']['assembly']['frac']))"; done
