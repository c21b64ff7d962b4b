# tests (subset by -k) + bench + fused-dprnn timeline
bash tools/gpu_tb.sh "$1" "$2"
bash tools/gpu_dfdebug.sh 2>&1 | head -6
