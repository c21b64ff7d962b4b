# usage: bash tools/gpu_ncu_train.sh <tag>: ncu --set full of a few launches of the largest training-step kernels -> raw CSV
tag=$1
for k in wgrad_kernel dw_bwd_weight_kernel gemm_tf32_kernel dw_bwd_data_kernel sru_scan_bwd_kernel; do
  timeout 400 ncu --set full --clock-control none -k "regex:$k" -s 8 -c 3 -o /tmp/${tag}_$k python tools/prof_train.py 1 > gpurun_out/${tag}_$k.log 2>&1; echo "$k ncu exit $?"
  ncu -i /tmp/${tag}_$k.ncu-rep --page raw --csv > gpurun_out/${tag}_${k}_raw.csv 2>/dev/null
done
