# final profile of a round: GPU tests, launch list, full ncu capture of the hot kernels, full bench line, reference arm
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
TAG=${TAG:-r1_final}
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gpu_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_gpu_tests.log
tail -4 gpurun_out/${TAG}_gpu_tests.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/${TAG}_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/${TAG}_launches_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'response_pass|assemble' -s 9 -c 3 -f -o gpurun_out/${TAG}_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
timeout 600 python bench.py --steps 200 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2>> gpurun_out/${TAG}_bench.err
tail -c 900 gpurun_out/${TAG}_bench.json
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
PLF_MG=1 timeout 900 python scripts/solve_demo.py 1024 5 0.004 2>&1 | grep -E "^N=|linear solves|increments|global|launches" > gpurun_out/${TAG}_solve_demo_1024.txt; cat gpurun_out/${TAG}_solve_demo_1024.txt | head -3
