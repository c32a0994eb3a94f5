mkdir -p gpurun_out; rm -f gpurun_out/bench_r34.log
run() { echo -n "$1 " >> gpurun_out/bench_r34.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r34.log 2>&1; }
export SDSP_B200_IIR_PIPE=0
run pipe=0 iir16384_f32 5; run pipe=0 iir18944_f32 5
cat gpurun_out/bench_r34.log
