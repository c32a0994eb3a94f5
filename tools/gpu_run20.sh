mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus2.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 4 > gpurun_out/bench_2gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2gpu.log
tail -c 2500 gpurun_out/bench_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 3 --warmup 1 --impl reference > gpurun_out/bench_2gpu_ref.log 2>&1; echo "rc=$?" >> gpurun_out/bench_2gpu_ref.log
tail -c 1200 gpurun_out/bench_2gpu_ref.log
timeout 300 python -m pytest tests/test_gpu_fft.py -m gpu -q --timeout 300 -k "large or all_sizes" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --workload fft65536_f32 | tail -c 900
