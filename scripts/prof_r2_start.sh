# Round-2 start profile at HEAD: config lines, ncu of every hot kernel (incl. SVC / Barlat), solve demo breakdown.
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${TAG:-r2_start}
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for c in j2_256 hill_512 "svc_512 16384" "barlat_2048 262144"; do timeout 300 python scripts/config_bench.py $c; done > gpurun_out/${TAG}_configs.jsonl 2>&1
cat gpurun_out/${TAG}_configs.jsonl | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'response_pass|assemble' -s 9 -c 3 -f -o gpurun_out/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'response_svc' -c 1 -f -o gpurun_out/${TAG}_svc python scripts/config_bench.py svc_512 4096 > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'response_model' -c 1 -f -o gpurun_out/${TAG}_barlat python scripts/config_bench.py barlat_2048 262144 > /dev/null 2>&1
PLF_BREAKDOWN=1 timeout 600 python scripts/solve_demo.py 1024 5 0.004 2>&1 | grep -E "^N=|breakdown|linear solves|increments|global|launches" > gpurun_out/${TAG}_solve_demo_1024.txt; cat gpurun_out/${TAG}_solve_demo_1024.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/${TAG}_solve256_launches.csv python scripts/solve_demo.py 256 5 0.004 > gpurun_out/${TAG}_solve256.log 2>&1
tail -3 gpurun_out/${TAG}_solve256.log
ls -la gpurun_out | tail -12
