timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_solve_ll --launch-skip 80 -c 1 -o gpurun_out/rl3_solve_s8 -f python bench.py --no-cpu-baseline --no-adapter --steps 20 --warmup 5 --groups 1 > gpurun_out/rl3_ncu.log 2>&1
tail -3 gpurun_out/rl3_ncu.log
ls -la gpurun_out/*.ncu-rep
