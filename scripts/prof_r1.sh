set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r1_launches_bench.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'response_hill|assemble_rows' -s 9 -c 3 -o gpurun_out/r1_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r1_prof_bench.log 2>&1
ls -la gpurun_out
python bench.py --steps 50 --warmup 5 --compact 0 --no-e2e --no-cpu 2>&1 | tail -1 > gpurun_out/r1_bench_nocompact.json
