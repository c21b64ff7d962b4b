# usage: bash tools/gpu_ncu_full.sh <out-name> <kernel-regex> <count> [skip]
set -x
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:$2" -s ${4:-0} -c $3 -o gpurun_out/$1 python tools/prof_forward.py 1 > gpurun_out/$1.log 2>&1; echo "ncu exit $?"
tail -3 gpurun_out/$1.log
