# compute-sanitizer memcheck over the GPU test-suite (bounded: the big-mesh tests are left out, memcheck runs kernels 10-50x slower)
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout ${T1:-330} compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_model.py -q -x -m gpu -k "not two_gpus and not scale" > gpurun_out/memcheck_model.log 2>&1
echo "model: exit $?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|at plf" gpurun_out/memcheck_model.log | head -12
timeout ${T2:-270} compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "not config4 and not config5 and not fast_mode and not full_size and not 1024" > gpurun_out/memcheck_parity.log 2>&1
echo "parity: exit $?"; grep -E "passed|failed|ERROR SUMMARY|Invalid|at plf" gpurun_out/memcheck_parity.log | head -12
