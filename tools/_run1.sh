mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_iir.py -m gpu -q --timeout 600 -x > gpurun_out/r02_pytest_iir.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_pytest_iir.log
tail -n 15 gpurun_out/r02_pytest_iir.log
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/r02_bench_iir_delta.log 2>&1; }
rm -f gpurun_out/r02_bench_iir_delta.log
for w in iir16384_f32 iir16384_f32_scan iir16384_f64 iir4096_f32_scan iirscan_f64 iirscan_f32; do run $w 5; done
SDSP_B200_IIR_PACK=1 run iir16384_f32 5
cat gpurun_out/r02_bench_iir_delta.log
