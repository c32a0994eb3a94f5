mkdir -p gpurun_out
L=gpurun_out/r02_fft_r2c.log
run() { timeout 300 python bench.py --steps $2 --warmup 3 $4 --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('e2e'))" >> $L 2>&1; }
rm -f $L
timeout 900 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 300 -x -k "half_spectrum or real_input" 2>&1 | tail -15
for rep in 1 2; do
run fftr2c4096_f32 20 "r2c" --no-e2e
run fftr2c4096_f64 20 "r2c" --no-e2e
run fftr2c65536_f32 20 "r2c" --no-e2e
run fftreal65536_f32 20 "full" --no-e2e
run fft16384_f32 20 "plain" --no-e2e
SDSP_B200_FFT_R32_ALIAS=1 run fft16384_f32 20 "alias" --no-e2e
done
run fftr2c4096_f32 6 "r2c-e2e"
cat $L
