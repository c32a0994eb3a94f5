python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
bash tools/gpu_validate.sh
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],4), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/r02_bench_workloads_v1.txt 2>&1; }
rm -f gpurun_out/r02_bench_workloads_v1.txt
for w in fft64_f32 fft256_f32 fft1024_f32 fft2048_f32 fft4096_f32 fft8192_f32 fft16384_f32 fft32768_f32 fft65536_f32 fft131072_f32 fft262144_f32 fft256_f64 fft1024_f64 fft2048_f64 fft4096_f64 iir16384_f32 iir16384_f32_scan iir16384_f64 iir18944_f32 iir4096_f32 iir4096_f32_lookback iirscan_f64 iirscan_f32 iirscan_f64_lookback pipeline65536_f32 pipeline_cfg5_f32; do run $w 10 x; done
SDSP_B200_FFT_FUSED_SMALL=0 run fft16384_f32 10 "single-cta"
cat gpurun_out/r02_bench_workloads_v1.txt
