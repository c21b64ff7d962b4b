# usage: bash tools/gpu_dftile.sh <tag>: GPU parity tests, bench with 128-position RNN tiles (default) and with RTFS_DF_TILE=256, phase timeline
tag=$1
rm -f gpurun_out/parity_report.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"; tail -6 gpurun_out/${tag}_tests.log
cp gpurun_out/parity_report.txt gpurun_out/${tag}_parity_report.txt 2>/dev/null
timeout 600 python bench.py --no-cpu --no-eager --no-train > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
RTFS_DF_TILE=256 timeout 600 python bench.py --no-cpu --no-eager --no-train > gpurun_out/${tag}_bench_t256.json 2>> gpurun_out/${tag}_bench.err
python - <<PY
import json
for f in ("${tag}_bench", "${tag}_bench_t256"):
    d = json.load(open("gpurun_out/" + f + ".json"))
    print(f, round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "dprnn", d["stages"]["dprnn_fused"], "roofline", d["roofline"]["frac"])
PY
RTFS_DF_DEBUG=1 python tools/prof_forward.py 1 2>&1 | grep dprnn_fused | head -12
