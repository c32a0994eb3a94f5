mkdir -p gpurun_out; rm -f gpurun_out/bench_r43.log
run() { echo -n "$1 " >> gpurun_out/bench_r43.log; timeout 200 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r43.log 2>&1; }
for t in 0 4 5; do SDSP_B200_FFT_TUNE=$t run tune=$t fft4096_f64 50; done
cat gpurun_out/bench_r43.log
SDSP_B200_FFT_TUNE=4 timeout 300 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 200 -k "all_sizes or f64" 2>&1 | tail -3
