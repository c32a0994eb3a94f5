mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_reference_tests.py -m gpu -q -s --timeout 600 > gpurun_out/r02_reference_tests_through_dropin.log 2>&1; echo "rc=$?" >> gpurun_out/r02_reference_tests_through_dropin.log
grep -E "benchmark|failed|passed|rc=" gpurun_out/r02_reference_tests_through_dropin.log
timeout 600 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 600 -x -k "all_sizes or fused or large_frames or host_buffers or real_input" 2>&1 | tail -3
timeout 300 python -m pytest tests/test_gpu_iir.py -m gpu -q --timeout 600 -x -k "more_sections or golden" 2>&1 | tail -3
