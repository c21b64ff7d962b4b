# usage: bash tools/gpu_ncu_csv.sh <out-name> <kernel-regex> <count> [skip]
# full ncu capture on the box, exported to CSV there (the .ncu-rep is dropped when > 40 MB so gpurun_out stays small)
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${4:-0} -c $3 -o /tmp/$1 python tools/prof_forward.py 1 > gpurun_out/$1.log 2>&1; echo "ncu exit $?"
ncu -i /tmp/$1.ncu-rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
ncu -i /tmp/$1.ncu-rep --page details --csv > gpurun_out/$1_details.csv 2>/dev/null
sz=$(stat -c %s /tmp/$1.ncu-rep); echo "rep size $sz"
if [ "$sz" -lt 40000000 ]; then cp /tmp/$1.ncu-rep gpurun_out/; fi
ls -la gpurun_out | tail -5
