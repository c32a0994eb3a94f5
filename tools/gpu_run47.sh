mkdir -p gpurun_out; rm -f gpurun_out/bench_r47.log
run() { echo -n "$1 " >> gpurun_out/bench_r47.log; timeout 200 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['config']['plan'][100:200])" >> gpurun_out/bench_r47.log 2>&1; }
run x fft8192_f32 30
cat gpurun_out/bench_r47.log
timeout 300 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 200 -k "all_sizes" 2>&1 | tail -2
