# One-GPU validation on a B200 box (gpurun -- bash tools/gpu_validate.sh): GPU tests, default bench line, reference arm, every workload.
mkdir -p gpurun_out; rm -f gpurun_out/bench_r44.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -n 5 gpurun_out/pytest_all.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; python - <<'PY'
import json
for l in open('gpurun_out/bench_default.log'):
    if l.startswith('{'):
        d=json.loads(l); print('default', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), 'cpu', round(d['cpu_baseline']['value']), [(s['workload'], round(s['value']), round(s['roofline']['frac'],4)) for s in d['secondary']], d['clocks'])
PY
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; tail -c 400 gpurun_out/bench_reference.log
run() { echo -n "$1 " >> gpurun_out/bench_r44.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/bench_r44.log 2>&1; }
for w in fft4096_f32 fft4096_f64 fft1024_f32 fft32768_f32 fft65536_f32 iir16384_f32 iir16384_f32_scan iir16384_f64 iir4096_f32_scan iirscan_f64 iirscan_f32 iirscan_f64_lookback pipeline65536_f32; do run x $w 5; done
cat gpurun_out/bench_r44.log
