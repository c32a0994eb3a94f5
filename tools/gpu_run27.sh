mkdir -p gpurun_out; rm -f gpurun_out/bench_r27.log
run() { echo -n "$1 " >> gpurun_out/bench_r27.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['config']['plan'][-215:-100])" >> gpurun_out/bench_r27.log 2>&1; }
for t in 0 6 7 8; do export SDSP_B200_SEG_TUNE=$t; run tune=$t iir4096_f32_scan 3; run tune=$t iirscan_f64 5; done
unset SDSP_B200_SEG_TUNE
cat gpurun_out/bench_r27.log
BI="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary --workload iir4096_f32_scan"
timeout 300 $BI > gpurun_out/plain_split.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_tma_kernel -s 6 -c 1 -o gpurun_out/prof_iir_split_f32_v1 $BI > gpurun_out/ncu_split.log 2>&1
tail -2 gpurun_out/ncu_split.log
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft4096_f32"
timeout 300 $BI > gpurun_out/plain_fft2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_cta_kernel -s 3 -c 1 -o gpurun_out/prof_fft4096_f32_v2 $BI > gpurun_out/ncu_fft2.log 2>&1
tail -2 gpurun_out/ncu_fft2.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/launches_fft4096_f32_v2.csv $BI > /dev/null 2>&1
