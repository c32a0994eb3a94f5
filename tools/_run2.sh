mkdir -p gpurun_out
run() { timeout 300 python bench.py --steps $2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $1 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$3', d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['clocks']['sm_mhz'], d['clocks']['reasons'])" >> gpurun_out/r02_iir_tune.log 2>&1; }
rm -f gpurun_out/r02_iir_tune.log
for rep in 1 2; do
for t in 0 1 2 4 7; do
SDSP_B200_TMA_TUNE=$t run iir16384_f32 5 "tune=$t"
done
for t in 0 2 5 1; do
SDSP_B200_TMA_TUNE=$t run iir16384_f64 5 "tune=$t"
done
done
cat gpurun_out/r02_iir_tune.log
BI="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary --workload iir16384_f32"
ncu --set full --clock-control none --import-source on -k regex:iir_tma_kernel -s 3 -c 1 -o gpurun_out/prof_iir16384_f32_r02 $BI > gpurun_out/ncu_b.log 2>&1
ls -la gpurun_out/*.ncu-rep
