mkdir -p gpurun_out; rm -f gpurun_out/bench_r42.log
timeout 600 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 200 -k "large or real_input or all_sizes" > gpurun_out/pytest_fft.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fft.log
tail -n 5 gpurun_out/pytest_fft.log
run() { echo -n "$1 " >> gpurun_out/bench_r42.log; timeout 200 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r42.log 2>&1; }
run remap fft65536_f32 20; run remap pipeline65536_f32 5
cat gpurun_out/bench_r42.log
