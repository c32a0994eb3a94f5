mkdir -p gpurun_out
L=gpurun_out/r02_fft_final2.log
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> $L 2>&1; }
rm -f $L
timeout 900 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 300 -x -k "edges or real_input or half_spectr or shutdown or host_buffers or data_mover" 2>&1 | tail -4
for rep in 1 2; do
run fftreal65536_f32 20 "lib"
run fftr2c65536_f32 20 "lib"
run fft65536_f32 20 "lib"
run fft32768_f32 20 "lib"
run pipeline_cfg5_f32 5 "lib"
run pipeline_cfg5_r2c_f32 5 "lib"
done
cat $L
