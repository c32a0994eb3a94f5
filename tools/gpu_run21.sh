mkdir -p gpurun_out; rm -f gpurun_out/bench_r21.log
timeout 900 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 300 > gpurun_out/pytest_fft.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_fft.log
tail -n 15 gpurun_out/pytest_fft.log
run() { echo -n "$1 " >> gpurun_out/bench_r21.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['config']['plan'][-90:])" >> gpurun_out/bench_r21.log 2>&1; }
run x fft65536_f32 20
run x pipeline65536_f32 5
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --workload pipeline65536_f32 > gpurun_out/bench_pipeline.log 2>&1
cat gpurun_out/bench_r21.log; tail -c 1500 gpurun_out/bench_pipeline.log
