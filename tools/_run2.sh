mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/r02_iir_b2one.log 2>&1; }
rm -f gpurun_out/r02_iir_b2one.log
for rep in 1 2; do
for m in 0 1; do
for w in iir16384_f32 iir16384_f32_scan iir4096_f32 iir16384_f64 iirscan_f64 iir18944_f32; do
SDSP_B200_IIR_B2ONE=$m run $w 5 "b2one=$m"
done; done; done
cat gpurun_out/r02_iir_b2one.log
