# builds the micro-benchmarks of tools/probe for sm_100a (run from the repo root; run the binaries through gpurun)
set -e
cd "$(dirname "$0")"
F="-O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a"
for p in umma_probe stride_probe inflight_probe inflight_probe2 cpasync_probe producer_probe dfgemm_probe dfissue_probe; do nvcc $F -o $p $p.cu; done
nvcc $F -o tcp_ablate tcp_ablate.cu                                   # production kernel, one ingredient removed at a time
nvcc $F -DRTFS_PROBE_W_ONCE -o tcp_ablate_w1 tcp_ablate.cu            # ... without the weight stream
nvcc $F -DRTFS_PROBE_NO_STAGE -o tcp_ablate_ns tcp_ablate.cu          # ... without the epilogue staging
nvcc $F -DRTFS_PROBE_TIMING -DREPS=1 -DWARM=1 -o tcp_ablate_t tcp_ablate.cu  # ... with per-role wait-cycle counters
