mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/r02_fft_tw_m16.log 2>&1; }
rm -f gpurun_out/r02_fft_tw_m16.log
for rep in 1 2; do
for lib in lib lib_m16; do
for w in fft4096_f32 fft4096_f64 fft256_f32 fft256_f64 fft1024_f32 fft65536_f32; do
SDSP_B200_LIB=$PWD/simpledsp_b200/$lib/libsdsp_b200.so run $w 30 "$lib"
done; done; done
cat gpurun_out/r02_fft_tw_m16.log
timeout 600 python -m pytest tests/test_gpu_fft.py tests/test_gpu_reference_tests.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
SDSP_B200_LIB=$PWD/simpledsp_b200/lib_m16/libsdsp_b200.so timeout 600 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 600 -x 2>&1 | tail -3
