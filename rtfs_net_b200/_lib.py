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
}


def parse_enum(name, header_text=None):
    """Enumerator names of `enum <name> { ... }` in the header, in order."""
    if header_text is None:
        with open(HEADER) as f:
            header_text = f.read()
    text = re.sub(r"/\*.*?\*/", "", header_text, flags=re.S)
    m = re.search(r"enum\s+" + name + r"\s*\{(.*?)\}", text, re.S)
    if m is None:
        raise RuntimeError(f"enum {name} not found in {HEADER}")
    body = m.group(1)
    names = []
    for tok in body.split(","):
        tok = tok.strip()
        if tok:
            names.append(tok.split("=")[0].strip())
    return names


def declared_functions(header_text=None):
    """Names of all functions the header declares."""
    if header_text is None:
        with open(HEADER) as f:
            header_text = f.read()
    text = re.sub(r"/\*.*?\*/", "", header_text, flags=re.S)
    return re.findall(r"\b(rtfs_[a-z0-9_]+)\s*\(", text)


PARAM_NAMES = parse_enum("rtfs_param")[:-1]  # drop RTFS_P_COUNT
WS_NAMES = parse_enum("rtfs_ws")[:-1]
STAT_NAMES = parse_enum("rtfs_stat")[:-1]
STAGE_NAMES = parse_enum("rtfs_stage")[:-1]
P = {n: i for i, n in enumerate(PARAM_NAMES)}
WS = {n: i for i, n in enumerate(WS_NAMES)}
ST = {n: i for i, n in enumerate(STAT_NAMES)}

ABI_VERSION = 5
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
