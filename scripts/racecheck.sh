cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout ${T1:-150} compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_model.py -q -x -m gpu -k "fused and 64" > gpurun_out/racecheck_mg.log 2>&1
echo "mg: exit $?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|at plf" gpurun_out/racecheck_mg.log | head -10
timeout ${T2:-150} compute-sanitizer --tool racecheck --print-limit 5 python -m pytest tests/test_gpu_parity.py -q -x -m gpu -k "mixed_materials or step_host_packed or strip_sharded" > gpurun_out/racecheck_resp.log 2>&1
echo "resp: exit $?"; grep -E "passed|failed|RACECHECK SUMMARY|hazard|at plf" gpurun_out/racecheck_resp.log | head -10
