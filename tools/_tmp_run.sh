mkdir -p gpurun_out
B="python bench.py --steps 20 --warmup 3 --no-e2e --no-cpu --no-secondary --workload"
P='import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["ms_per_step"], d["value"], d["roofline"]["frac"], d["self_check"])'
for pr in 128 256 0; do
export SDSP_B200_FFT_FUSED_PROMO=$pr
for w in fft65536_f32 fft32768_f32; do
    echo "== promo $pr $w"; timeout 60 $B $w | python -c "$P"
done
done
