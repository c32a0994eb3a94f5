bash tools/gpu_validate.sh
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench_default.csv $B > gpurun_out/ncu_launches.log 2>&1
ls -la gpurun_out/r02_launches_bench_default.csv
