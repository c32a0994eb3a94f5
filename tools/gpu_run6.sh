mkdir -p gpurun_out; rm -f gpurun_out/tune_iir.log
for promo in 128 256; do for cfg in 0 1 2 3 4 5 6 7 8; do
  export SDSP_B200_IIR_TUNE=$cfg,$promo
  echo -n "cfg=$cfg promo=$promo : " >> gpurun_out/tune_iir.log
  timeout 200 python bench.py --steps 3 --warmup 3 --no-e2e --no-cpu --workload iir16384_f32 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],3), round(d['roofline']['achieved'],1), round(d['roofline']['frac'],4))" >> gpurun_out/tune_iir.log 2>&1
done; done
unset SDSP_B200_IIR_TUNE
cat gpurun_out/tune_iir.log
