mkdir -p gpurun_out
export SDSP_B200_IIR_PIPE=1
BI="python bench.py --steps 1 --warmup 3 --no-e2e --no-cpu --no-secondary --workload iir16384_f32"
timeout 300 $BI > gpurun_out/plain_pipe.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:iir_tma_pipe_kernel -s 3 -c 1 -o gpurun_out/prof_iir_pipe_f32_v1 $BI > gpurun_out/ncu_pipe.log 2>&1
tail -2 gpurun_out/ncu_pipe.log
