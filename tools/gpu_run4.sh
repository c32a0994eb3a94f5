mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_iir.py -m gpu -q -x --timeout 120 > gpurun_out/pytest_iir.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_iir.log
tail -n 5 gpurun_out/pytest_iir.log
if grep -q "pytest rc=0" gpurun_out/pytest_iir.log; then
  BI="python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload iir16384_f32"
  timeout 300 $BI > gpurun_out/bench_iir_tma2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_tma_kernel -s 1 -c 1 -o gpurun_out/prof_iir_tma2_f32 $BI > gpurun_out/ncu_full_iir_tma2.log 2>&1
  tail -n 2 gpurun_out/bench_iir_tma2.log | cut -c1-400
fi
