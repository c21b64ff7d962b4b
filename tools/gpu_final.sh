# usage: bash tools/gpu_final.sh <tag>: GPU tests, smoke, default bench (with the cpu_baseline leg) and the ncu launch list
tag=$1
rm -f gpurun_out/parity_report.txt
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"; tail -1 gpurun_out/${tag}_tests.log
cp gpurun_out/parity_report.txt gpurun_out/${tag}_parity_report.txt
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:stft|gemm|caf|dwroll|dprnn|att|rowblock" -s 177 -c 130 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list exit $?"
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"block frac",round(d["roofline_block"]["frac"],4),"ms_per_pass",round(d["roofline_block"]["ms_per_pass"],3))
print(d["roofline"]["kernel"], round(d["roofline"]["frac"],3), d["roofline_hbm"]["kernel"], round(d["roofline_hbm"]["frac"],3), d["clocks"])
PY
