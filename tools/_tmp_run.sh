mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_fft.py tests/test_gpu_reference_tests.py -x -q 2>&1 | tail -3
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-secondary --workload"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["self_check"])'
for w in fft65536_f32 fft32768_f32 pipeline65536_f32; do
    echo "== $w"; timeout 60 $B $w | python -c "$P"
done
BI="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-secondary --workload fft65536_f32"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:fft_fused_tma_kernel -s 3 -c 1 -o gpurun_out/prof_fft65536_f32_fused_tma_v2 $BI > gpurun_out/ncu_f.log 2>&1
tail -n 1 gpurun_out/ncu_f.log
