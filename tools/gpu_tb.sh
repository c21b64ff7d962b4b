# usage: bash tools/gpu_tb.sh <tag> [pytest -k expr]   -- GPU parity tests + short bench, outputs under gpurun_out/
tag=$1; kexpr=$2
if [ -n "$kexpr" ]; then
  timeout 600 python -m pytest tests -m gpu -x -q -k "$kexpr" > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"
else
  timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"
fi
tail -25 gpurun_out/${tag}_tests.log
timeout 600 python bench.py --no-cpu > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
tail -5 gpurun_out/${tag}_bench.err
python - <<PY
import json
try:
    d=json.load(open("gpurun_out/${tag}_bench.json"))
    print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"block frac",round(d["roofline_block"]["frac"],4),"ms_per_pass",round(d["roofline_block"]["ms_per_pass"],3))
    for k,v in d["stages"].items(): print(f"  {k:18s} {v['ms_per_launch']:8.4f} x{v['launches_per_step']:4.0f} = {v['ms_per_step']:8.4f}  {v.get('gbps',0):8.1f} GB/s")
except Exception as e: print("no bench json", e)
PY
