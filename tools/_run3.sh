mkdir -p gpurun_out
L=gpurun_out/r02_fft_r32.log
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> $L 2>&1; }
rm -f $L
timeout 900 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 300 -x -k "all_sizes or 8192 or 16384" 2>&1 | tail -4
for rep in 1 2; do
for w in fft8192_f32 fft16384_f32; do
SDSP_B200_FFT_R32=0 run $w 20 "four-pass"
run $w 20 "radix32"
done; done
cat $L
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft16384_f32 --frames 4096"
ncu --set full --clock-control none --import-source on -k regex:fft_cta -s 2 -c 1 -o gpurun_out/prof_fft16384_f32_r32 $BI > gpurun_out/ncu_r.log 2>&1
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft8192_f32 --frames 8192"
ncu --set full --clock-control none --import-source on -k regex:fft_cta -s 2 -c 1 -o gpurun_out/prof_fft8192_f32_r32 $BI > gpurun_out/ncu_r2.log 2>&1
ls -la gpurun_out/*.ncu-rep
