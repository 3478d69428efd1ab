set -x
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'response_pass|assemble_rows' -s 9 -c 3 -o gpurun_out/r1b_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > gpurun_out/r1b_prof_bench.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r1b_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
python bench.py --steps 100 --warmup 5 --no-cpu 2>&1 | tail -1 > gpurun_out/r1b_bench.json
