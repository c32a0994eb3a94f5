mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 300 > gpurun_out/pytest_all.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_all.log
tail -n 6 gpurun_out/pytest_all.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -4
