"""TEST INFRASTRUCTURE ONLY -- CPU oracle for the SRU layer (never imported by the product path).

Restates one layer of the third-party `sru.SRU` stack used by the reference's DualPathRNN
(/root/reference/src/models/layers/rnn_layers.py:100-105,150).  The `sru` source is not in
/root/reference (pinned `sru==2.6.0` / git HEAD, setup/requirements.yaml:18,33); the recurrence
below follows the published algorithm as written down in SURVEY.md App. C.
PARITY UNPINNED for this component.

Two interchangeable scans:
  * `oracle/_build/libsru_scan.so` (plain C + OpenMP, oracle/sru_scan.c) -- used when built and no
    gradient is required (this is what the CPU baseline times, mirroring upstream's C++ CPU loop);
  * a torch loop over time (vectorised over columns) -- differentiable, dtype-generic.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libsru_scan.so")
_lib = None


def _load_lib():
    global _lib
    if _lib is None and os.path.exists(_LIB_PATH):
        lib = ctypes.CDLL(_LIB_PATH)
        for name in ("sru_scan_f32", "sru_scan_f64"):
            fn = getattr(lib, name)
            fn.restype = None
            fn.argtypes = [ctypes.c_void_p] * 6 + [ctypes.c_int] * 5
        _lib = lib
    return _lib


def sru_scan_torch(U, x, weight_c, bias, d, ndir, k):
    """U (L,B,D,k), x (L,B,D) -> h (L,B,D), c_last (B,D).  Differentiable."""
    L, B, D, _ = U.shape
    vf, vr = weight_c[:D], weight_c[D:]
    bf, br = bias[:D], bias[D:]
    hs = [None] * L
    # forward half uses time order, backward half reversed order: run both as one loop over s
    rev = torch.zeros(D, dtype=torch.bool, device=U.device)
    if ndir == 2:
        rev[d:] = True
    c = U.new_zeros(B, D)
    idx_f = torch.arange(L, device=U.device)
    # gather per-step rows: for column j at step s the time index is s (fwd) or L-1-s (bwd)
    Uf = U
    Ub = U.flip(0)
    xf = x
    xb = x.flip(0) if x is not None else None
    h_steps = []
    for s in range(L):
        u = torch.where(rev[None, :, None], Ub[s], Uf[s])  # (B,D,k)
        f = torch.sigmoid(u[..., 1] + vf * c + bf)
        r = torch.sigmoid(u[..., 2] + vr * c + br)
        if k == 4:
            xp = u[..., 3]
        else:
            xp = torch.where(rev[None, :], xb[s], xf[s])
        c = f * c + (1.0 - f) * u[..., 0]
        h_steps.append(r * c + (1.0 - r) * xp)
    hsf = torch.stack(h_steps)  # indexed by scan step
    h = torch.where(rev[None, None, :], hsf.flip(0), hsf)
    return h, c


def sru_layer_forward(x, weight, weight_c, bias, hidden_size, bidirectional):
    """One SRU layer.  x (L,B,in) time-major -> (h (L,B,D), c_last (B,D))."""
    L, B, n_in = x.shape
    ndir = 2 if bidirectional else 1
    D = hidden_size * ndir
    k = weight.shape[1] // D
    assert k in (3, 4) and (k == 4 or n_in == D)
    U = (x.reshape(L * B, n_in) @ weight).view(L, B, D, k)
    need_grad = torch.is_grad_enabled() and (x.requires_grad or weight.requires_grad)
    lib = _load_lib()
    if lib is not None and not need_grad and x.device.type == "cpu" and x.dtype in (torch.float32, torch.float64):
        U = U.contiguous()
        xc = x.contiguous()
        wc = weight_c.detach().contiguous().to(x.dtype)
        bs = bias.detach().contiguous().to(x.dtype)
        h = torch.empty(L, B, D, dtype=x.dtype)
        c = torch.empty(B, D, dtype=x.dtype)
        fn = lib.sru_scan_f32 if x.dtype == torch.float32 else lib.sru_scan_f64
        fn(U.data_ptr(), xc.data_ptr(), wc.data_ptr(), bs.data_ptr(), h.data_ptr(), c.data_ptr(), L, B, hidden_size, ndir, k)
        return h, c
    return sru_scan_torch(U, x if k == 3 else None, weight_c, bias, hidden_size, ndir, k)
