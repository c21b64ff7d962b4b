# usage: bash tools/gpu_quick.sh <tag>: GPU parity tests + inference bench line (no baselines / training block) + A/B of the video fork
tag=$1
rm -f gpurun_out/parity_report.txt
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests exit $?"; tail -4 gpurun_out/${tag}_tests.log
cp gpurun_out/parity_report.txt gpurun_out/${tag}_parity_report.txt 2>/dev/null
grep "video block" gpurun_out/parity_report.txt
timeout 600 python bench.py --no-cpu --no-eager --no-train > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench exit $?"
RTFS_NO_VIDEO_FORK=1 timeout 600 python bench.py --no-cpu --no-eager --no-train > gpurun_out/${tag}_bench_nofork.json 2>> gpurun_out/${tag}_bench.err
python - <<PY
import json
for f in ("${tag}_bench", "${tag}_bench_nofork"):
    d = json.load(open("gpurun_out/" + f + ".json"))
    print(f, round(d["value"], 1), round(d["ms_per_step"], 3), "e2e", round(d["e2e"]["value"], 1), "launches", d["gpu_launches"])
PY
