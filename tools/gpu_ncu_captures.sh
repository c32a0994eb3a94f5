# ncu captures on one B200 (gpurun -- bash tools/gpu_ncu_captures.sh); summaries: python tools/ncu_summary.py <rep> > profiles/<name>.txt
mkdir -p gpurun_out
# launch list of the default bench command (every BASELINE config), short run
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu"
timeout 600 $B > gpurun_out/plain_default.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches_bench_default.csv $B > gpurun_out/ncu_launches.log 2>&1
# the headline kernel (config 2 fp32) at full size
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft4096_f32"
timeout 300 $BI > gpurun_out/plain_h.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_cta_kernel -s 3 -c 1 -o gpurun_out/prof_fft4096_f32_r02 $BI > gpurun_out/ncu_h.log 2>&1
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft65536_f32 --frames 1024"
timeout 300 $BI > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_fused_tma2 -s 2 -c 1 -o gpurun_out/prof_fft65536_f32_r02 $BI > gpurun_out/ncu_a.log 2>&1
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fftreal65536_f32 --frames 2048"
timeout 300 $BI > gpurun_out/plain_r.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_real64k -s 2 -c 1 -o gpurun_out/prof_fftreal65536_f32_r02b $BI > gpurun_out/ncu_r.log 2>&1
ls -la gpurun_out/*.ncu-rep gpurun_out/*.csv
