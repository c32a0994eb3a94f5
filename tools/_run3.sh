mkdir -p gpurun_out
L=gpurun_out/r02_fft_alias_pf.log
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> $L 2>&1; }
rm -f $L
for rep in 1 2; do
for lib in lib lib_noap; do
for w in fft4096_f64 fft1024_f64 fft2048_f64 fft256_f64; do
SDSP_B200_LIB=$PWD/simpledsp_b200/$lib/libsdsp_b200.so run $w 20 "$lib"
done; done; done
cat $L
