mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fft.py -m gpu -q -x --timeout 300 > gpurun_out/pytest_fft.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fft.log
tail -n 25 gpurun_out/pytest_fft.log
rm -f gpurun_out/bench_r8.log
for w in fft65536_f32 fft4096_f32; do
timeout 300 python bench.py --steps 10 --warmup 3 --no-e2e --no-cpu --workload $w 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['gpu_launches'])" >> gpurun_out/bench_r8.log 2>&1
done
cat gpurun_out/bench_r8.log
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload iirscan_f64"
timeout 300 $BI > gpurun_out/plain_scan.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_scan_kernel -s 1 -c 1 -o gpurun_out/prof_iir_scan_f64 $BI > gpurun_out/ncu_scan.log 2>&1
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload iir16384_f32"
timeout 300 $BI > gpurun_out/plain_tma3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_tma_kernel -s 1 -c 1 -o gpurun_out/prof_iir_tma3_f32 $BI > gpurun_out/ncu_tma3.log 2>&1
ls gpurun_out/*.ncu-rep
