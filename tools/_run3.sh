mkdir -p gpurun_out
L=gpurun_out/r02_fft_staged.log
run() { timeout 300 python bench.py --steps $2 --warmup 3 $4 --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d.get('e2e'))" >> $L 2>&1; }
rm -f $L
timeout 900 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 300 -x -k "staged_16384 or all_sizes or host_buffers" 2>&1 | tail -15
for rep in 1 2; do
run fft16384_f32 20 "staged" --no-e2e
SDSP_B200_FFT_STAGED=0 run fft16384_f32 20 "plain" --no-e2e
done
run pipeline_cfg5_r2c_f32 5 "r2c" --no-e2e
run pipeline_cfg5_f32 5 "full" --no-e2e
cat $L
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft16384_f32 --frames 4096"
ncu --set full --clock-control none --import-source on -k regex:fft_cta -s 2 -c 1 -o gpurun_out/prof_fft16384_f32_staged $BI > gpurun_out/ncu_r.log 2>&1
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fftr2c4096_f32 --frames 16384"
ncu --set full --clock-control none --import-source on -k regex:fft_r2c -s 2 -c 1 -o gpurun_out/prof_fftr2c4096_f32 $BI > gpurun_out/ncu_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep
