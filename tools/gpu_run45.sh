mkdir -p gpurun_out; rm -f gpurun_out/bench_r45.log
run() { echo -n "$1 " >> gpurun_out/bench_r45.log; timeout 200 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r45.log 2>&1; }
run plain fft4096_f32 100; run plain fft4096_f64 50; run plain fft65536_f32 20
export SDSP_B200_LIB=$PWD/simpledsp_b200/lib_cs/libsdsp_b200.so
run cs fft4096_f32 100; run cs fft4096_f64 50; run cs fft65536_f32 20
cat gpurun_out/bench_r45.log
