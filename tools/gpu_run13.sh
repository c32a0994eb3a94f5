mkdir -p gpurun_out; rm -f gpurun_out/bench_r13.log
timeout 900 python -m pytest tests/test_gpu_iir.py -m gpu -q --timeout 300 > gpurun_out/pytest_iir.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_iir.log
tail -n 30 gpurun_out/pytest_iir.log
run() { echo -n "$1 " >> gpurun_out/bench_r13.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['config']['plan'][-220:])" >> gpurun_out/bench_r13.log 2>&1; }
run x iirscan_f64 5; run x iirscan_f32 5; run x iir4096_f32_scan 3; run x iir16384_f64 5
for rows in 9472 37888 75776; do export SDSP_B200_SEG_ROWS=$rows; run rows=$rows iirscan_f64 5; run rows=$rows iir4096_f32_scan 3; done
cat gpurun_out/bench_r13.log
