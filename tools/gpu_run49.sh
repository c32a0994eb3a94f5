mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_iir.py -m gpu -q --timeout 300 -k "two_devices or ragged or bit_identical" > gpurun_out/pytest_2dev.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_2dev.log; tail -15 gpurun_out/pytest_2dev.log
