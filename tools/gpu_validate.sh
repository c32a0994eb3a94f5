# One-GPU validation on a B200 box (gpurun -- bash tools/gpu_validate.sh): GPU tests, default bench line (every BASELINE config), reference arm.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -n 8 gpurun_out/pytest_all.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.err
python - <<'PY'
import json
for l in open('gpurun_out/bench_default.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('default', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), round(d['e2e']['pcie_GBps_each_way'],1), 'cpu', round(d['cpu_baseline']['value']), d['clocks'])
        for s in d['secondary']:
            print('  ', s['workload'], round(s['value']), round(s['ms_per_step'],3), round(s['roofline']['frac'],4), s['self_check'], s['clocks']['sm_mhz'], s['clocks']['reasons'], ('e2e %d Msamples/s %.1f GB/s each way' % (s['e2e']['value'], s['e2e']['pcie_GBps_each_way'])) if 'e2e' in s else '')
            print('      ', s['plan'][:230])
PY
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_reference.log 2>&1; tail -c 600 gpurun_out/bench_reference.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
