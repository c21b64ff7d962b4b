"""profiles/r0N_traffic.json from an `ncu --page raw --csv` export of tools/prof_forward.py: measured DRAM bytes per launch of
the kernel behind each stage (first matching launch).  usage: python tools/make_traffic.py raw.csv summary-name > profiles/r02_traffic.json"""
import csv, json, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {n: i for i, n in enumerate(hdr)}
def val(r, m):
    v = float(r[col[m]].replace(",", "")); u = units[col[m]].lower()
    return v * {"byte": 1.0, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9, "ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u, 1.0)
stages = [("RTFS_SG_STFT", "stft16_kernel"), ("RTFS_SG_ENC_CONV", "Im2colLoader3x"), ("RTFS_SG_BOTTLENECK", "GlnActLoader<256"), ("RTFS_SG_GATE_PROJ", "GateLoader"),
          ("RTFS_SG_DW_S1", "dwroll_kernel<XrGln<2>, 1, 0, 288"), ("RTFS_SG_DW_S2_POOL", "dwroll_kernel<XrGln<0>, 1, 1, 288"), ("RTFS_SG_DPRNN_FUSED", "dprnn_fused_kernel<1, 128, 1>"),
          ("RTFS_SG_ATT_QKV", "att_conv_ln_tc_kernel<96"), ("RTFS_SG_ATT_CORE", "attn_core_tc_kernel"), ("RTFS_SG_ATT_PROJ", "att_conv_ln_tc_kernel<64"),
          ("RTFS_SG_TFAR_GLOBAL", "dwroll_kernel<XrPlain, 4"), ("RTFS_SG_TFAR_CAT_GLOBAL", "dwroll_kernel<XrTfarP, 2"), ("RTFS_SG_TFAR_CAT_LOCAL", "dwroll_kernel<XrTfarP, 1"),
          ("RTFS_SG_RESID_OUT", "TfarLoader, ResidOutEpi4"), ("RTFS_SG_RESID_OUT_CAF", "TfarLoader, ResidOutCafEpi4"), ("RTFS_SG_MASK_DEC", "MaskDecEpi4"), ("RTFS_SG_VIDEO", "video_block_kernel"), ("RTFS_SG_DEC_ISTFT", "dec_istft16_kernel")]
out = {"source": "ncu --set full --clock-control none, python tools/prof_forward.py 1 (B=32, 2 s, RTFS-Net-4), first launch of each kernel; " + sys.argv[2], "stages": {}}
for st, pat in stages:
    for r in data:
        if pat in r[col["Kernel Name"]]:
            rd, wr = val(r, "dram__bytes_read.sum"), val(r, "dram__bytes_write.sum")
            out["stages"][st] = {"kernel": r[col["Kernel Name"]][:90], "dram_bytes": rd + wr, "dram_read_bytes": rd, "dram_write_bytes": wr, "ncu_duration_us": val(r, "gpu__time_duration.sum")}
            break
print(json.dumps(out, indent=1))
