mkdir -p gpurun_out; rm -f gpurun_out/bench_r33.log
run() { echo -n "$1 " >> gpurun_out/bench_r33.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r33.log 2>&1; }
for p in 0 1; do export SDSP_B200_IIR_PIPE=$p; run pipe=$p iir16384_f32 5; run pipe=$p iir16384_f32_pitch 5; done
cat gpurun_out/bench_r33.log
