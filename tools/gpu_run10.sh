mkdir -p gpurun_out; rm -f gpurun_out/bench_r10.log
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -n 8 gpurun_out/pytest_all.log
for tune in 0 1; do for w in iirscan_f64 iir4096_f32_scan; do
export SDSP_B200_SCAN_TUNE=$tune
echo -n "tune=$tune " >> gpurun_out/bench_r10.log
timeout 300 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload $w 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'])" >> gpurun_out/bench_r10.log 2>&1
done; done
unset SDSP_B200_SCAN_TUNE
cat gpurun_out/bench_r10.log
export SDSP_B200_SCAN_TUNE=1
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload iirscan_f64"
timeout 300 $BI > gpurun_out/plain_scan.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_scan_kernel -s 1 -c 1 -o gpurun_out/prof_iir_scan3_f64 $BI > gpurun_out/ncu_scan.log 2>&1
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload iir4096_f32_scan"
timeout 300 $BI > gpurun_out/plain_scan32.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_scan_kernel -s 1 -c 1 -o gpurun_out/prof_iir_scan3_f32 $BI > gpurun_out/ncu_scan32.log 2>&1
