mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference_tests.py -m gpu -q -s > gpurun_out/reftests.log 2>&1; echo "rc=$?" >> gpurun_out/reftests.log
B="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu"
$B > gpurun_out/plain_f32.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_fft_f32.csv $B > gpurun_out/ncu_launch.log 2>&1
$B > gpurun_out/plain_f32b.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_cta_kernel -s 3 -c 2 -o gpurun_out/prof_fft_f32 $B > gpurun_out/ncu_full_f32.log 2>&1
B64="python bench.py --steps 4 --warmup 3 --no-e2e --no-cpu --workload fft4096_f64"
$B64 > gpurun_out/plain_f64.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:fft_cta_kernel -s 3 -c 1 -o gpurun_out/prof_fft_f64 $B64 > gpurun_out/ncu_full_f64.log 2>&1
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload iir16384_f32"
$BI > gpurun_out/plain_iir.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_seq_kernel -s 1 -c 1 -o gpurun_out/prof_iir_f32 $BI > gpurun_out/ncu_full_iir.log 2>&1
tail -n 12 gpurun_out/reftests.log
tail -n 3 gpurun_out/ncu_launch.log gpurun_out/ncu_full_f32.log gpurun_out/ncu_full_f64.log gpurun_out/ncu_full_iir.log
ls -la gpurun_out
