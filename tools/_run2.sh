mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'], d['details']['plan'][130:200])" >> gpurun_out/r02_iir_seg_rows.log 2>&1; }
rm -f gpurun_out/r02_iir_seg_rows.log
for rep in 1 2; do
for r in 0 32768 65536 98304 131072 163840 262144; do
SDSP_B200_SEG_ROWS=$r run iir4096_f32 5 "rows=$r"
done
for r in 0 131072 262144 393216 524288; do
SDSP_B200_SEG_ROWS=$r run iir16384_f32_scan 5 "rows=$r"
done
done
cat gpurun_out/r02_iir_seg_rows.log
