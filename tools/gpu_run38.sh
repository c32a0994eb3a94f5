mkdir -p gpurun_out; rm -f gpurun_out/bench_r38.log
timeout 600 python -m pytest tests/test_gpu_iir.py -m gpu -q --timeout 300 > gpurun_out/pytest_iir.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_iir.log
tail -n 4 gpurun_out/pytest_iir.log
run() { echo -n "$1 " >> gpurun_out/bench_r38.log; timeout 300 python bench.py --steps $3 --warmup 3 --no-e2e --no-cpu --no-secondary --workload $2 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['config']['workload'], round(d['ms_per_step'],3), round(d['value']), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4), d['self_check'], d['config']['plan'][-200:-90])" >> gpurun_out/bench_r38.log 2>&1; }
for w in iir16384_f32_scan iir4096_f32_scan iirscan_f64 iirscan_f32; do run x $w 5; done
SDSP_B200_IIR_PACK=0 run scalar iir18944_f32 5
SDSP_B200_IIR_PACK=1 run packed iir18944_f32 5
cat gpurun_out/bench_r38.log
