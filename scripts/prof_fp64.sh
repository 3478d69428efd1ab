# FP64 flop per element update of every config: thread-level DFMA / DADD / DMUL counts of the return-mapping kernels
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
M=smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,gpu__time_duration.sum
run() { local name=$1; shift; timeout 400 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/fp64_${name}.csv "$@" > /dev/null 2>&1; }
run target -k regex:'response_pass' -s 6 -c 2 python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-configs
run c3 -k regex:'response_pass' -s 6 -c 2 python bench.py --config c3 --steps 1
run c4 -k regex:'response_svc' -c 1 python scripts/config_bench.py svc_512 8192
run c5 -k regex:'response_model' -c 2 python scripts/config_bench.py barlat_2048 262144
head -5 gpurun_out/fp64_c5.csv | cut -c1-300
