"""Cuts the launches of forwards [first, last] (1-based; a forward starts at its stft16_kernel launch) out of an ncu launch-list CSV.
usage: launch_slice.py list.csv first last > slice.csv"""
import sys
lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
hdr, rows = lines[0], lines[1:]
first, last = int(sys.argv[2]), int(sys.argv[3])
seg = 0
out = []
for l in rows:
    if '"stft16_kernel' in l:  # not dec_istft16_kernel
        seg += 1
    if first <= seg <= last:
        out.append(l)
sys.stdout.write(hdr)
sys.stdout.writelines(out)
