set -x
python tools/grid_gpu.py --out gpurun_out/grid_gpu_final.npz > gpurun_out/r02r_grid_gpu.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/r02r_bench_1gpu.json 2> gpurun_out/r02r_bench.err; tail -3 gpurun_out/r02r_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02r_bench_ref.json 2>> gpurun_out/r02r_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02r_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-parity --no-single-runs --no-config5 --no-strong > gpurun_out/r02r_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pr_ensemble_kernel -c 1 -f -o gpurun_out/r02r_headline python tools/run_headline.py --members 8192 --order desc > gpurun_out/r02r_ncu.log 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:pr_ensemble_kernel -c 1 --csv --log-file gpurun_out/r02r_traffic_full_launch.csv python tools/run_headline.py --members 65536 --order desc > /dev/null 2>&1
PR_LONG_POLL=1 ncu --set full --clock-control none --import-source on -k regex:pr_long_fused -s 3 -c 1 -f -o gpurun_out/r02r_long_fused python tools/bench_long.py --repeat 1 --check 0 --members 256 > gpurun_out/r02r_ncu_long.log 2>&1
ls -la gpurun_out/ | tail -12
