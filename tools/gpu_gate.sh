for g in 0 1 2; do
  echo "=== RTFS_DF_GATE=$g"
  rm -f gpurun_out/parity_report.txt
  RTFS_DF_GATE=$g timeout 300 python -m pytest tests -m gpu -x -q -k "dprnn or full_forward or block_full" 2>&1 | tail -1
  grep -E "dprnn|forward\[|out " gpurun_out/parity_report.txt | head -12
  RTFS_DF_GATE=$g timeout 300 python bench.py --no-cpu --steps 10 --warmup 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',round(d['value'],1), 'dprnn', d['stages']['dprnn_fused'])"
done
