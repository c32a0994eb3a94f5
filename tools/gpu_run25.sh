mkdir -p gpurun_out
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft65536_f32 --frames 1024"
timeout 300 $BI > gpurun_out/plain_f64k.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_fused64k -s 2 -c 1 -o gpurun_out/prof_fft_fused64k_v1 $BI > gpurun_out/ncu_f64k.log 2>&1
tail -3 gpurun_out/ncu_f64k.log
