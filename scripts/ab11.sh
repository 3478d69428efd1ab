cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
for c in j2_256 hill_512 "svc_512 16384" "barlat_2048 262144"; do timeout 300 python scripts/config_bench.py $c 2>&1 | tail -1 | tee -a gpurun_out/config_bench_s3.jsonl; done
