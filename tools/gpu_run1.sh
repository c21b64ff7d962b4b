set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv
python tools/l2bw.py > gpurun_out/l2bw.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r01b_tests.log 2>&1; echo "tests exit $?"
tail -3 gpurun_out/r01b_tests.log
timeout 600 python bench.py > gpurun_out/r01b_bench.json 2> gpurun_out/r01b_bench.err; echo "bench exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:rtfs -c 46 -o gpurun_out/r01b_full python tools/prof_forward.py 1 > gpurun_out/r01b_ncu.log 2>&1; echo "ncu exit $?"
ls -la gpurun_out
