mkdir -p gpurun_out
for w in iirscan_f64 iirscan_f32 iir4096_f32_scan; do for t in 0 5; do
export SDSP_B200_SEG_TUNE=$t
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_${w}_t$t.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --workload $w > gpurun_out/ncu_ll.log 2>&1
done; done
ls -la gpurun_out/launches_iir*
