# N-GPU bench line (gpurun --gpus N -- bash tools/gpu_multi_gpu_bench.sh): the default line with every BASELINE config, config 5 strong-scaled;
# then the 2-rank GPU tests when at least two GPUs are there.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus=$N" > gpurun_out/gpus.txt
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.log 2> gpurun_out/bench_${N}gpu.err; echo "rc=$?"
tail -3 gpurun_out/bench_${N}gpu.err
python - <<PY
import json
for l in open('gpurun_out/bench_${N}gpu.log'):
    if l.startswith('{'):
        d=json.loads(l)
        print('N=%d default' % d['n_gpus'], round(d['value']), round(d['ms_per_step'],3), round(d['roofline']['frac'],4), 'e2e', round(d['e2e']['value']), d['clocks']['reasons'])
        for s in d['secondary']:
            print('  ', s['workload'], s['scaling'], round(s['value']), round(s['ms_per_step'],3), round(s['roofline']['frac'],4), s['self_check'], ('e2e %d' % s['e2e']['value']) if 'e2e' in s else '')
PY
if [ "$N" -ge 2 ]; then timeout 600 python -m pytest tests/test_gpu_multi_rank.py tests/test_gpu_iir.py -m gpu -q --timeout 300 -k "multi or two_devices" 2>&1 | tail -3; fi
