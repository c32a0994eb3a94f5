mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_iir.py -m gpu -q --timeout 600 -x 2>&1 | tail -4
timeout 600 python -m pytest tests/test_gpu_reference_tests.py -m gpu -q -s --timeout 600 > gpurun_out/r02_reference_tests_through_dropin.log 2>&1; echo "rc=$?" >> gpurun_out/r02_reference_tests_through_dropin.log
grep -E "benchmark|failed|passed|rc=" gpurun_out/r02_reference_tests_through_dropin.log
SDSP_B200_NO_CHAIN=1 timeout 600 python -m pytest tests/test_gpu_reference_tests.py -m gpu -q -s --timeout 600 2>&1 | grep -E "filter benchmark|passed"
