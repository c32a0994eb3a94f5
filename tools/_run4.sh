bash tools/gpu_validate.sh
bash tools/gpu_ncu_captures.sh
