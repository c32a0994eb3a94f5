mkdir -p gpurun_out; rm -f gpurun_out/bench_r23.log
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -n 5 gpurun_out/pytest_all.log
run() { echo -n "$1 " >> gpurun_out/bench_r23.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r23.log 2>&1; }
for pf in 1 0; do export SDSP_B200_FFT_PREFETCH=$pf; run pf=$pf fft4096_f32 100; run pf=$pf fft4096_f64 50; run pf=$pf fft1024_f32 50; done
unset SDSP_B200_FFT_PREFETCH
for t in 1 4; do SDSP_B200_FFT_TUNE=$t run pf=1,tune=$t fft4096_f32 100; done
cat gpurun_out/bench_r23.log
timeout 600 python bench.py > gpurun_out/bench_default.log 2>&1; tail -c 4000 gpurun_out/bench_default.log
