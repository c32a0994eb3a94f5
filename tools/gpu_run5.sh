mkdir -p gpurun_out
for v in "" _A _B _K _Jn _Bn; do
  export SDSP_B200_LIB=$PWD/simpledsp_b200/lib$v/libsdsp_b200.so
  echo "== variant '$v'" >> gpurun_out/variants.log
  timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload iir16384_f32 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['roofline']['achieved'], d['roofline']['frac'])" >> gpurun_out/variants.log 2>&1
done
unset SDSP_B200_LIB
timeout 300 python -m pytest tests/test_gpu_iir.py -m gpu -q -x --timeout 120 2>&1 | tail -3 >> gpurun_out/variants.log
cat gpurun_out/variants.log
