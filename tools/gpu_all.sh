# usage: bash tools/gpu_all.sh <tag>: every GPU test + inference bench line (no baselines / training block)
tag=$1
rm -f gpurun_out/parity_report.txt
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/${tag}_tests.log
cp gpurun_out/parity_report.txt gpurun_out/${tag}_parity_report.txt 2>/dev/null
timeout 600 python bench.py --no-cpu --no-eager --no-train > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
python - <<PY
import json
d = json.load(open("gpurun_out/${tag}_bench.json"))
print("${tag}", round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "roofline", round(d["roofline"]["frac"], 4), "block", round(d["roofline_block"]["frac"], 4))
for k, v in d["stages"].items(): print(f"  {k:18s} {v['ms_per_launch']:8.4f} x{v['launches_per_step']:4.0f} = {v['ms_per_step']:8.4f}  {v.get('gbps', '')}")
PY
