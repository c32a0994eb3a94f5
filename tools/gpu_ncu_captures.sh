# ncu --set full captures of the three headline kernels (gpurun -- bash tools/gpu_ncu_captures.sh); summaries: python tools/ncu_summary.py <rep> > profiles/<name>.txt
mkdir -p gpurun_out
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft65536_f32 --frames 1024"
timeout 300 $BI > gpurun_out/plain_a.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_fused64k -s 2 -c 1 -o gpurun_out/prof_fft_fused64k_v3 $BI > gpurun_out/ncu_a.log 2>&1
BI="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary --workload iir16384_f32"
timeout 300 $BI > gpurun_out/plain_b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_tma_kernel -s 3 -c 1 -o gpurun_out/prof_iir16384_f32_v5 $BI > gpurun_out/ncu_b.log 2>&1
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft4096_f64"
timeout 300 $BI > gpurun_out/plain_c.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_cta_kernel -s 3 -c 1 -o gpurun_out/prof_fft4096_f64_v2 $BI > gpurun_out/ncu_c.log 2>&1
