# usage: bash tools/gpu_round.sh <tag>: full GPU evidence run (tests, bench, ncu launch list, ncu full of the main kernels -> CSV)
tag=$1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/${tag}_smi.txt
rm -f gpurun_out/parity_report.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"; tail -2 gpurun_out/${tag}_tests.log
cp gpurun_out/parity_report.txt gpurun_out/${tag}_parity_report.txt
timeout 900 python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${tag}_bench_reference.json 2>> gpurun_out/${tag}_bench.err; echo "ref exit $?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "regex:stft|gemm|caf|dwroll|dprnn|att|rowblock" -s 177 -c 130 --csv --log-file gpurun_out/${tag}_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > /dev/null 2>&1; echo "ncu list exit $?"
timeout 1200 ncu --set full --clock-control none --import-source on -k "regex:gemm_tc|gemm_tf32|dprnn_fused|dwroll|attn_core|att_conv_ln|stft" -c 32 -o /tmp/${tag}_full python tools/prof_forward.py 1 > gpurun_out/${tag}_ncu_full.log 2>&1; echo "ncu full exit $?"
ncu -i /tmp/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full_raw.csv 2>/dev/null
ls -la /tmp/${tag}_full.ncu-rep
python - <<PY
import json
d=json.load(open("gpurun_out/${tag}_bench.json"))
print("value",round(d["value"],1),"ms/step",round(d["ms_per_step"],3),"e2e",round(d["e2e"]["value"],1),"block frac",round(d["roofline_block"]["frac"],4),"ms_per_pass",round(d["roofline_block"]["ms_per_pass"],3), "cpu", d.get("cpu_baseline"))
for k,v in d["stages"].items(): print(f"  {k:18s} {v['ms_per_launch']:8.4f} x{v['launches_per_step']:4.0f} = {v['ms_per_step']:8.4f}  {v.get('gbps',0):8.1f} GB/s")
PY
