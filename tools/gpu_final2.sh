# usage: bash tools/gpu_final2.sh <tag>: every GPU test, smoke, the default bench line (all legs), the reference arm, the ncu launch
# list of the two timed steps of `bench.py --steps 2 --warmup 3` and an ncu --set full capture of one forward's kernels (CSV)
tag=$1
rm -f gpurun_out/parity_report.txt gpurun_out/train_report.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"; tail -1 gpurun_out/${tag}_tests.log
cp gpurun_out/parity_report.txt gpurun_out/${tag}_parity_report.txt
cp gpurun_out/train_report.txt gpurun_out/${tag}_train_report.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "reference arm exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:stft|gemm|caf|dwroll|dprnn|att|video|istft" -c 900 --csv --log-file gpurun_out/${tag}_launches_all.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-eager --no-train > /dev/null 2>&1; echo "ncu list exit $?"
python tools/launch_slice.py gpurun_out/${tag}_launches_all.csv 5 6 > gpurun_out/${tag}_launches.csv
python tools/launch_summary.py gpurun_out/${tag}_launches.csv > gpurun_out/${tag}_launches_summary.txt; head -12 gpurun_out/${tag}_launches_summary.txt
rm -f gpurun_out/${tag}_launches_all.csv
timeout 1500 ncu --set full --clock-control none --import-source on -k "regex:stft|gemm|caf|dwroll|dprnn|att|video|istft" -c 64 -o /tmp/${tag}_full python tools/prof_forward.py 1 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/${tag}_full_raw.csv > gpurun_out/${tag}_ncu_full_summary.txt; head -5 gpurun_out/${tag}_ncu_full_summary.txt
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"block frac",round(d["roofline_block"]["frac"],4),"ms_per_pass",round(d["roofline_block"]["ms_per_pass"],3))
print(d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline_hbm"]["kernel"], round(d["roofline_hbm"]["frac"],3), d["clocks"])
print("train", json.dumps(d.get("train_step"))[:400])
print("eager", json.dumps(d.get("gpu_eager_baseline"))[:200])
print("cpu", json.dumps(d.get("cpu_baseline"))[:300])
PY
