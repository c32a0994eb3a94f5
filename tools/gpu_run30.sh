mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --print-limit 30 --error-exitcode 7 python -m pytest tests/test_gpu_iir.py tests/test_gpu_fft.py -m gpu -q --timeout 600 -x -k "ragged_shapes or few_channels or real_input or unaligned or single_long_channel" > gpurun_out/sanitizer_memcheck.log 2>&1; echo "rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -n 25 gpurun_out/sanitizer_memcheck.log
