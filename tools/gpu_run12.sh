mkdir -p gpurun_out; rm -f gpurun_out/bench_r12.log
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -n 8 gpurun_out/pytest_all.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 3000 gpurun_out/bench_default.log
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 1500 gpurun_out/bench_reference.log
run() { echo -n "$1 " >> gpurun_out/bench_r12.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r12.log 2>&1; }
run x iirscan_f64 3; run x iir4096_f32_scan 3; run x iir16384_f32 3; run x fft4096_f64 30; run x fft65536_f32 10; run x fft1024_f32 30
cat gpurun_out/bench_r12.log
