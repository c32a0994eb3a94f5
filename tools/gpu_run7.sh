mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_iir.py -m gpu -q -x --timeout 200 > gpurun_out/pytest_iir.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_iir.log
tail -n 30 gpurun_out/pytest_iir.log
if grep -q "pytest rc=0" gpurun_out/pytest_iir.log; then
  for w in iirscan_f64 iir4096_f32_scan iir16384_f32 iir4096_f32; do
  timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload $w 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_scan.log 2>&1
  done
  cat gpurun_out/bench_scan.log
fi
