mkdir -p gpurun_out; rm -f gpurun_out/bench_r11.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -n 6 gpurun_out/pytest_all.log
run() { echo -n "$1 " >> gpurun_out/bench_r11.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r11.log 2>&1; }
for tune in 0 1; do export SDSP_B200_SCAN_TUNE=$tune; run "scan_tune=$tune" iirscan_f64 3; run "scan_tune=$tune" iir4096_f32_scan 3; done
unset SDSP_B200_SCAN_TUNE
for tune in 0 1 2 3 4; do export SDSP_B200_FFT_TUNE=$tune; run "fft_tune=$tune" fft4096_f32 50; done
for tune in 0 1 2 3; do export SDSP_B200_FFT_TUNE=$tune; run "fft_tune=$tune" fft4096_f64 30; done
unset SDSP_B200_FFT_TUNE
cat gpurun_out/bench_r11.log
