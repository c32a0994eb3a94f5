mkdir -p gpurun_out; rm -f gpurun_out/bench_r35.log
timeout 900 python -m pytest tests/test_gpu_iir.py tests/test_gpu_reference_tests.py -m gpu -q --timeout 300 > gpurun_out/pytest_iir.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_iir.log
tail -n 12 gpurun_out/pytest_iir.log
run() { echo -n "$1 " >> gpurun_out/bench_r35.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r35.log 2>&1; }
for p in 0 1; do export SDSP_B200_IIR_PIPE=$p; run pipe=$p iir16384_f32 5; done
unset SDSP_B200_IIR_PIPE
run x iir16384_f32_scan 5; run x iir4096_f32_scan 5; run x iirscan_f32 5; run x iir18944_f32 5
cat gpurun_out/bench_r35.log
