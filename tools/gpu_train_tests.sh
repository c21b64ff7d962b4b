# usage: bash tools/gpu_train_tests.sh <tag> [pytest -k expression]: the training-step GPU tests with the full report
tag=$1
rm -f gpurun_out/train_report.txt
timeout 1500 python -m pytest tests/test_gpu_train.py -m gpu -q ${2:+-k "$2"} > gpurun_out/${tag}_train_tests.log 2>&1; echo "train tests exit $?"
tail -25 gpurun_out/${tag}_train_tests.log
cp gpurun_out/train_report.txt gpurun_out/${tag}_train_report.txt 2>/dev/null
