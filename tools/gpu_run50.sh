mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_multi_rank.py -m gpu -q --timeout 300 > gpurun_out/pytest_nccl.log 2>&1; echo "rc=$?" >> gpurun_out/pytest_nccl.log; tail -15 gpurun_out/pytest_nccl.log
