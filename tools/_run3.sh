mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_fft.py tests/test_gpu_reference_tests.py -m gpu -q --timeout 300 -x -k "half_spectr or additions" 2>&1 | tail -8
