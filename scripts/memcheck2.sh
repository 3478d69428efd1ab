cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout ${T1:-300} compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_model.py -q -x -m gpu -k "umat or svc_with or MLparam or tresca or sdim3 or interleaved or fused or calc_scf or averages" > gpurun_out/memcheck_model2.log 2>&1
echo "model2: exit $?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|at plf" gpurun_out/memcheck_model2.log | head -12; tail -2 gpurun_out/memcheck_model2.log | cut -c1-200
