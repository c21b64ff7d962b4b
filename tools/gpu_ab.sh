# usage: bash tools/gpu_ab.sh "<ENV=.. ENV2=..>" ... : quick bench (stage table for the GEMM stages) per environment setting
for e in "$@"; do
  echo "=== $e"
  env $e timeout 600 python bench.py --no-cpu --steps 10 --warmup 3 > /tmp/ab.json 2>/tmp/ab.err || tail -3 /tmp/ab.err
  python - <<PY
import json
d=json.load(open("/tmp/ab.json"))
print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3))
for k,v in d["stages"].items(): print(f"  {k:18s} {v['ms_per_launch']:8.4f} x{v['launches_per_step']:4.0f} = {v['ms_per_step']:8.4f}")
PY
done
