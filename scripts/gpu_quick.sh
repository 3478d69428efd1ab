cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6
for c in "svc_512 16384" "barlat_2048 262144" "barlat_2048"; do timeout 300 python scripts/config_bench.py $c 2>&1 | cut -c1-330; done
