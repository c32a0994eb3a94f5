mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_default.log 2> gpurun_out/bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/bench_default.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('default', round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['clocks'])
        for s in d['secondary']:
            print('  ', s['workload'], round(s['value']), round(s['ms_per_step'],3), round(s['roofline']['frac'],4), s['steps'], s['step_ms'], s['clocks']['sm_mhz'], s['clocks']['reasons'])
PY
