mkdir -p gpurun_out
SDSP_B200_FFT_FUSED_TMA=2 timeout 600 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 300 -x -k "all_sizes or large_frames or fused_65536 or real_input or host_buffers" 2>&1 | tail -5
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['details']['plan'][-60:])" >> gpurun_out/r02_fft_tune.log 2>&1; }
rm -f gpurun_out/r02_fft_tune.log
for rep in 1 2; do
for m in 1 2; do
for w in fft65536_f32 fft32768_f32 fft131072_f32 fft262144_f32 pipeline65536_f32; do
SDSP_B200_FFT_FUSED_TMA=$m run $w 20 "mode=$m"
done; done; done
cat gpurun_out/r02_fft_tune.log
