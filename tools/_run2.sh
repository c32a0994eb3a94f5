mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/r02_iir_rows_tune.log 2>&1; }
rm -f gpurun_out/r02_iir_rows_tune.log
for rep in 1 2; do
for t in 0 1 2 3 4 5 7; do
SDSP_B200_SEG_TUNE=$t run iirscan_f64 10 "seg_tune=$t"
SDSP_B200_SEG_TUNE=$t run iir4096_f32 5 "seg_tune=$t"
done; done
cat gpurun_out/r02_iir_rows_tune.log
