# N-GPU weak-scaling bench lines (gpurun --gpus 8 -- bash tools/gpu_multi_gpu_bench.sh): default workload and the config-5 pipeline.
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus=$N" > gpurun_out/gpus8.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 50 --warmup 4 > gpurun_out/bench_${N}gpu.log 2>&1; echo "rc=$?" >> gpurun_out/bench_${N}gpu.log
tail -c 1800 gpurun_out/bench_${N}gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus $N --steps 5 --warmup 3 --no-e2e --workload pipeline65536_f32 > gpurun_out/bench_${N}gpu_pipeline.log 2>&1; echo "rc=$?" >> gpurun_out/bench_${N}gpu_pipeline.log
tail -c 900 gpurun_out/bench_${N}gpu_pipeline.log
