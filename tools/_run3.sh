mkdir -p gpurun_out
PYTHONPATH=$PWD timeout 600 python tools/experiments/r2c_sweep.py > gpurun_out/r02_r2c_sweep.log 2>&1; cat gpurun_out/r02_r2c_sweep.log
