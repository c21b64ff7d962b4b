# usage: bash tools/gpu_train_bench.sh <tag>: bench line with the train_step block + kernel launch list of one training step
tag=$1
timeout 900 python bench.py --no-cpu --no-eager --steps 5 --warmup 3 > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("value",round(d["value"],1)); print(json.dumps(d.get("train_step"),indent=1))
PY
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${tag}_train_launches.csv python tools/prof_train.py > gpurun_out/${tag}_prof.log 2>&1; echo "ncu exit $?"
python tools/launch_summary.py gpurun_out/${tag}_train_launches.csv > gpurun_out/${tag}_train_launches_summary.txt 2>&1; head -45 gpurun_out/${tag}_train_launches_summary.txt
