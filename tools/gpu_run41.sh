mkdir -p gpurun_out; rm -f gpurun_out/bench_r41.log
timeout 600 python -m pytest tests/test_gpu_fft.py tests/test_gpu_reference_tests.py -m gpu -q --timeout 200 > gpurun_out/pytest_fft.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fft.log
tail -n 5 gpurun_out/pytest_fft.log
run() { echo -n "$1 " >> gpurun_out/bench_r41.log; timeout 200 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r41.log 2>&1; }
run packed fft4096_f32 100; run packed fft1024_f32 50; run packed fft65536_f32 20; run packed pipeline65536_f32 5
cat gpurun_out/bench_r41.log
