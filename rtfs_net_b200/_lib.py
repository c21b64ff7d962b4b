"""ctypes binding of librtfs_b200.so (the C ABI declared in include/rtfs_b200.h).

There is no fallback: if the library is missing or does not export a declared symbol the import
of the compute path raises, and every call that returns non-zero raises RuntimeError with the
library's error text.
"""
import ctypes
import os
import re

HERE = os.path.dirname(os.path.abspath(__file__))
HEADER = os.path.join(os.path.dirname(HERE), "include", "rtfs_b200.h")
LIB_PATH = os.environ.get("RTFS_B200_LIB", os.path.join(HERE, "lib", "librtfs_b200.so"))

_c_p = ctypes.c_void_p
_c_i = ctypes.c_int
_c_ll = ctypes.c_longlong
_c_f = ctypes.c_float

# name -> (restype, argtypes); must list every function include/rtfs_b200.h declares
PROTOTYPES = {
    "rtfs_abi_version": (_c_i, []),
    "rtfs_last_error": (ctypes.c_char_p, []),
    "rtfs_last_launch_count": (_c_ll, []),
    "rtfs_profile_enable": (None, [_c_i]),
    "rtfs_profile_collect": (_c_i, [ctypes.POINTER(ctypes.c_float), ctypes.POINTER(_c_i)]),
    "rtfs_ws_plan": (_c_ll, [_c_i, _c_i, _c_i, ctypes.POINTER(_c_ll)]),
    "rtfs_encoder_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_bottleneck_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_block_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_dprnn_forward": (_c_i, [_c_p, _c_i, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_mhsa_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_caf_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_p]),
    "rtfs_mask_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_decoder_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_avnet_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "rtfs_avnet_forward_av": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_p, _c_p]),
    "rtfs_video_forward": (_c_i, [_c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_mouth_preprocess": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_f, _c_f, _c_p]),
    "rtfs_wav_normalize": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_f, _c_p]),
    "rtfs_video_pack_plan": (_c_i, [ctypes.POINTER(_c_i), ctypes.POINTER(_c_i)]),
    # training step
    "rtfs_train_plan": (_c_ll, [_c_i, _c_i, _c_i, _c_i, ctypes.POINTER(_c_ll), ctypes.POINTER(_c_ll), ctypes.POINTER(_c_ll), ctypes.POINTER(_c_ll)]),
    "rtfs_train_pass_plan": (_c_ll, [_c_i, _c_i, ctypes.POINTER(_c_ll)]),
    "rtfs_avnet_train_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_i, _c_i, _c_i, _c_p]),
    "rtfs_avnet_backward": (_c_i, [_c_p] * 11 + [_c_i] * 5 + [_c_p]),
    "rtfs_block_train_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_block_backward": (_c_i, [_c_p] * 7 + [_c_i, _c_i, _c_p]),
    "rtfs_dprnn_train_forward": (_c_i, [_c_p, _c_i, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_dprnn_backward": (_c_i, [_c_p, _c_p, _c_i] + [_c_p] * 5 + [_c_i, _c_i, _c_p]),
    "rtfs_mhsa_train_forward": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_i, _c_i, _c_p]),
    "rtfs_mhsa_backward": (_c_i, [_c_p] * 7 + [_c_i, _c_i, _c_p]),
    "rtfs_snr_loss": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_p, _c_i, _c_i, ctypes.c_float, _c_p]),
    "rtfs_adamw_step": (_c_i, [_c_p, _c_p, _c_p, _c_p, _c_ll] + [ctypes.c_float] * 5 + [_c_i, ctypes.c_float, ctypes.c_float, _c_p, _c_p]),
}


def parse_enum(name, header_text):
    """Enumerator names of `enum <name> { ... }` in the header text, in order."""
    text = re.sub(r"/\*.*?\*/", "", header_text, flags=re.S)
    m = re.search(r"enum\s+" + name + r"\s*\{(.*?)\}", text, re.S)
    if m is None:
        raise RuntimeError(f"enum {name} not found in the header")
    names = []
    for tok in m.group(1).split(","):
        tok = tok.strip()
        if tok:
            names.append(tok.split("=")[0].strip())
    return names


def header_tables(header_text=None):
    """(params, ws, stats, stages, functions, abi_version) as include/rtfs_b200.h declares them.  Used by the build
    (which freezes them into _abi.py so that the installed package does not need the repo's include/ directory) and by
    the tests (which check that _abi.py is in sync)."""
    if header_text is None:
        with open(HEADER) as f:
            header_text = f.read()
    text = re.sub(r"/\*.*?\*/", "", header_text, flags=re.S)
    fns = sorted(set(re.findall(r"\b(rtfs_[a-z0-9_]+)\s*\(", text)))
    ver = int(re.search(r"#define\s+RTFS_ABI_VERSION\s+(\d+)", header_text).group(1))
    return (parse_enum("rtfs_param", header_text)[:-1], parse_enum("rtfs_ws", header_text)[:-1], parse_enum("rtfs_stat", header_text)[:-1],
            parse_enum("rtfs_stage", header_text)[:-1], fns, ver, parse_enum("rtfs_tape", header_text)[:-1], parse_enum("rtfs_bwd", header_text)[:-1])


def write_abi(path=None):
    """Regenerate rtfs_net_b200/_abi.py from the header (called by rtfs_net_b200.build)."""
    params, ws, stats, stages, fns, ver, tape, bwd = header_tables()
    path = path or os.path.join(HERE, "_abi.py")
    body = ['"""GENERATED by rtfs_net_b200.build from include/rtfs_b200.h -- do not edit.\n\nEnum tables of the C ABI, frozen into the package so that it works when copied / installed without the repo root."""',
            f"ABI_VERSION = {ver}"]
    for name, vals in (("PARAM_NAMES", params), ("WS_NAMES", ws), ("STAT_NAMES", stats), ("STAGE_NAMES", stages), ("FUNCTIONS", fns),
                       ("TAPE_NAMES", tape), ("BWD_NAMES", bwd)):
        body.append(f"{name} = [\n" + "".join(f'    "{v}",\n' for v in vals) + "]")
    text = "\n".join(body) + "\n"
    if not os.path.exists(path) or open(path).read() != text:
        with open(path, "w") as f:
            f.write(text)
    return path


def declared_functions(header_text=None):
    """Names of all functions the C ABI declares (from the header when it is present, else the frozen table)."""
    if header_text is None and not os.path.exists(HEADER):
        return list(_abi.FUNCTIONS)
    return header_tables(header_text)[4]


from . import _abi  # noqa: E402  (generated; see write_abi)

PARAM_NAMES = list(_abi.PARAM_NAMES)
WS_NAMES = list(_abi.WS_NAMES)
STAT_NAMES = list(_abi.STAT_NAMES)
STAGE_NAMES = list(_abi.STAGE_NAMES)
TAPE_NAMES = list(getattr(_abi, "TAPE_NAMES", []))
BWD_NAMES = list(getattr(_abi, "BWD_NAMES", []))
P = {n: i for i, n in enumerate(PARAM_NAMES)}
WS = {n: i for i, n in enumerate(WS_NAMES)}
ST = {n: i for i, n in enumerate(STAT_NAMES)}

ABI_VERSION = _abi.ABI_VERSION
_lib = None


def lib():
    """The loaded library (raises if it is missing: there is no CPU / eager fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m rtfs_net_b200.build` "
                "(the RTFS-Net B200 path has no fallback implementation)"
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)  # AttributeError if the symbol is not exported
            fn.restype = res
            fn.argtypes = args
        if handle.rtfs_abi_version() != ABI_VERSION:
            raise RuntimeError("librtfs_b200.so ABI version mismatch")
        _lib = handle
    return _lib


def check(code, what):
    if code != 0:
        msg = lib().rtfs_last_error()
        raise RuntimeError(f"{what} failed ({code}): {msg.decode() if msg else '?'}")


def ws_plan(B, L, Tv):
    """(total_bytes, {WS name: byte offset})"""
    offs = (_c_ll * len(WS_NAMES))()
    total = lib().rtfs_ws_plan(int(B), int(L), int(Tv), offs)
    return int(total), {n: int(offs[i]) for i, n in enumerate(WS_NAMES)}


def profile_enable(on=True):
    lib().rtfs_profile_enable(1 if on else 0)


def profile_collect():
    """{stage name: (total ms, launches)} since the last collect; synchronises the device."""
    n = len(STAGE_NAMES)
    ms = (ctypes.c_float * n)()
    cnt = (_c_i * n)()
    check(lib().rtfs_profile_collect(ms, cnt), "rtfs_profile_collect")
    return {STAGE_NAMES[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}


def train_plan(B, L, Tv, repeats):
    """Training tape / backward scratch layout: dict(tape_bytes, tape_offsets, pass_bytes, pass_offsets, bwd_bytes, bwd_offsets)."""
    to = (_c_ll * len(TAPE_NAMES))()
    bo = (_c_ll * len(BWD_NAMES))()
    pb, bb = _c_ll(0), _c_ll(0)
    total = lib().rtfs_train_plan(int(B), int(L), int(Tv), int(repeats), to, ctypes.byref(pb), ctypes.byref(bb), bo)
    po = (_c_ll * len(WS_NAMES))()
    lib().rtfs_train_pass_plan(int(B), int(L), po)
    return dict(tape_bytes=int(total), tape_offsets={n: int(to[i]) for i, n in enumerate(TAPE_NAMES)}, pass_bytes=int(pb.value),
                pass_offsets={n: int(po[i]) for i, n in enumerate(WS_NAMES)}, bwd_bytes=int(bb.value),
                bwd_offsets={n: int(bo[i]) for i, n in enumerate(BWD_NAMES)})


VIDEO_FIELDS = ["GW_W", "GW_B", "GW_A", "PJ_WT", "PJ_S", "PJ_T", "PJ_A", "DS_W", "DS_S", "DS_T", "LN1_G", "LN1_B", "PE", "IN_WT", "IN_B", "OUT_WT", "OUT_B",
                "LN2_G", "LN2_B", "F1_WT", "F1_G", "F1_B", "FR_W", "FR_B", "F2_WT", "F2_G", "F2_B", "TF", "RC_WT", "RC_B"]  # enum vp_field (csrc/video.cuh)


def video_pack_plan():
    """({field: float offset}, total floats) of the packed VP-block parameter buffer (RTFS_P_VIDEO_PACK)."""
    offs = (_c_i * 64)()
    n = _c_i(0)
    total = lib().rtfs_video_pack_plan(offs, ctypes.byref(n))
    if n.value != len(VIDEO_FIELDS):
        raise RuntimeError("video pack layout mismatch between _lib.VIDEO_FIELDS and csrc/video.cuh")
    return {f: int(offs[i]) for i, f in enumerate(VIDEO_FIELDS)}, int(total)
