mkdir -p gpurun_out; rm -f gpurun_out/bench_r37.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -n 6 gpurun_out/pytest_all.log
run() { echo -n "$1 " >> gpurun_out/bench_r37.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r37.log 2>&1; }
for w in iir16384_f32 iir18944_f32 iir16384_f32_scan iir16384_f64 iir4096_f32_scan iirscan_f64; do run x $w 5; done
cat gpurun_out/bench_r37.log
