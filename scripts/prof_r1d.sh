cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'response_pass|assemble' -s 9 -c 3 -o gpurun_out/r1d_prof python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r1d_launches.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
