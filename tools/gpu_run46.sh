mkdir -p gpurun_out; rm -f gpurun_out/bench_r46.log
run() { echo -n "$1 " >> gpurun_out/bench_r46.log; timeout 200 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['config']['plan'][30:200])" >> gpurun_out/bench_r46.log 2>&1; }
for w in fft64_f32 fft256_f32 fft1024_f32 fft2048_f32 fft8192_f32 fft16384_f32; do run x $w 30; done
cat gpurun_out/bench_r46.log
