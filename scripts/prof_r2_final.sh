# Round-2 profile at HEAD: launch list, ncu --set full of every hot kernel (target + configs 3/4/5 + solver) summarised on
# the box (the reports themselves are too large to bring back together), bench lines, reference arm
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${TAG:-r2_final}
NCU="ncu --set full --clock-control none --import-source on -f"
if [ -z "$SKIP_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; tail -3 gpurun_out/${TAG}_gpu_tests.log; fi
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs > /dev/null 2>&1
cap() {  # name, fp64 args..., then the command
  local name=$1; shift; local fp="$1"; shift
  timeout 500 $NCU "$@" > /dev/null 2>&1
  python scripts/summarize_ncu.py /tmp/${name}.ncu-rep gpurun_out/${TAG}_${name} $fp > /dev/null 2>&1
  rm -f /tmp/${name}.ncu-rep
}
cap target "--fp64 target 1048576" -k regex:'response_pass|assemble' -s 9 -c 3 -o /tmp/target python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs
cap c3 "--fp64 c3 262144" -k regex:'response_pass|assemble' -s 9 -c 3 -o /tmp/c3 python bench.py --config c3 --steps 1
cap c4 "--fp64 c4 8192" -k regex:'response_svc' -c 1 -o /tmp/c4 python scripts/config_bench.py svc_512 8192
cap c5 "--fp64 c5 262144" -k regex:'response_model' -c 2 -o /tmp/c5 python scripts/config_bench.py barlat_2048 262144
cap solver "" -k regex:'mg_jacobi|stencil_spmv|mg_coop|mg_residual' --launch-skip 3000 -c 12 -o /tmp/solver python scripts/solve_demo.py 1024 5 0.004
ls gpurun_out/ | head -30
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
tail -c 300 gpurun_out/${TAG}_bench.json; echo
du -sh gpurun_out
