RTFS_DF_DEBUG=1 python tools/prof_forward.py 1 2>&1 | grep dprnn_fused | head -12
