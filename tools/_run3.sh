mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 300 -x 2>&1 | tail -8
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/r02_fft_real64k.log 2>&1; }
for rep in 1; do
for m in 1 0; do
SDSP_B200_FFT_REAL64K=$m run pipeline_cfg5_f32 5 "real64k=$m"
done; done
cat gpurun_out/r02_fft_real64k.log
