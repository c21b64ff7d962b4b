# usage: bash tools/gpu_ncu_src.sh <out-name> <kernel-regex> <count> "<launch indices to export>" [skip]
# full ncu capture on the box; raw/details CSV of all captured launches plus the SASS source page (stall samples per
# instruction) of the listed launches are exported there, so nothing large travels back
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${5:-0} -c $3 -o /tmp/$1 python tools/prof_forward.py 1 > gpurun_out/$1.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
for i in $4; do
  ncu -i /tmp/$1.ncu-rep --page source --csv --print-source sass --launch-skip $i --launch-count 1 > gpurun_out/$1_src$i.csv 2>/dev/null
done
ls -la gpurun_out | tail -6
