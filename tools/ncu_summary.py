"""Summarise an `ncu --page raw --csv` export: one line per profiled launch with the metrics that matter here."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
want = [("gpu__time_duration.sum", "us", 1e-3), ("dram__bytes_read.sum", "rdMB", None), ("dram__bytes_write.sum", "wrMB", None),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", 1), ("lts__t_bytes.sum", "l2MB", None),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%", 1), ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", 1),
        ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tens%", 1), ("sm__inst_executed_pipe_uniform.sum", None, None),
        ("launch__registers_per_thread", "regs", 1), ("launch__occupancy_limit_shared_mem", "occSm", 1), ("launch__occupancy_limit_registers", "occRg", 1),
        ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bankc", 1), ("smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "stLong", 1),
        ("smsp__warp_issue_stalled_barrier_per_warp_active.pct", "stBar", 1), ("smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "stMath", 1),
        ("smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "stLG", 1), ("smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct", "stMIO", 1),
        ("smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct", "stShort", 1), ("smsp__warp_issue_stalled_wait_per_warp_active.pct", "stWait", 1),
        ("smsp__warp_issue_stalled_sleeping_per_warp_active.pct", "stSleep", 1), ("smsp__warp_issue_stalled_membar_per_warp_active.pct", "stMembar", 1)]
def num(s):
    try: return float(s.replace(",", ""))
    except Exception: return float("nan")
def tomb(v, u):
    u = u.lower()
    return v * {"byte": 1e-6, "kbyte": 1e-3, "mbyte": 1.0, "gbyte": 1e3}.get(u, 1e-6)
for r in data:
    name = r[col["Kernel Name"]][:70]
    out = [f"{r[col['ID']]:>3} {name:70s}"]
    for m, lab, sc in want:
        if lab is None or m not in col: continue
        v = num(r[col[m]]); u = units[col[m]]
        if lab.endswith("MB"): v = tomb(v, u)
        elif lab == "us": v = v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(u, 1e-3)
        out.append(f"{lab}={v:.1f}")
    print(" ".join(out))
