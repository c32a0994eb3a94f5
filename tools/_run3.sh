mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 600 -x 2>&1 | tail -5
