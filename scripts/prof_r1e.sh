# ncu full capture (with source) of the split kernels (staged pass A, FM3 pass B) and of the fused kernel
cd $GRAFT_REPO_ROOT; mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:'response_pass' -s 6 -c 2 -o gpurun_out/r1e_split -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --compact 1 --opt passa_direct=0 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'response_pass' -s 6 -c 2 -o gpurun_out/r1e_direct -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --compact 1 --opt passa_direct=1 > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:'response_fused' -s 3 -c 1 -o gpurun_out/r1e_fused -f python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --compact 2 --opt passa_direct=0 > /dev/null 2>&1
ls -la gpurun_out/*.ncu-rep
