# ncu captures on one B200 (gpurun -- bash tools/gpu_ncu_captures.sh); summaries: python tools/ncu_summary.py <rep> > profiles/<name>.txt
mkdir -p gpurun_out
# launch list of the default bench command (every BASELINE config), short run
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 600 $B > gpurun_out/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches_bench_default.csv $B > gpurun_out/ncu_launches.log 2>&1
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft65536_f32 --frames 1024"
timeout 300 $BI > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_fused_tma2 -s 2 -c 1 -o gpurun_out/prof_fft65536_f32_r02 $BI > gpurun_out/ncu_a.log 2>&1
BI="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary --workload iirscan_f64"
timeout 300 $BI > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_tma_kernel -s 6 -c 1 -o gpurun_out/prof_iirscan_f64_r02 $BI > gpurun_out/ncu_b.log 2>&1
BI="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary --workload iir4096_f32"
timeout 300 $BI > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_tma_kernel -s 6 -c 1 -o gpurun_out/prof_iir4096_f32_r02 $BI > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
