mkdir -p gpurun_out; rm -f gpurun_out/bench_r36.log
run() { echo -n "$1 " >> gpurun_out/bench_r36.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r36.log 2>&1; }
run packed iir16384_f32 5; run packed iir4096_f32_scan 5
export SDSP_B200_LIB=$PWD/simpledsp_b200/lib_np/libsdsp_b200.so
run nopack iir16384_f32 5; run nopack iir4096_f32_scan 5; run nopack iir18944_f32 5
cat gpurun_out/bench_r36.log
