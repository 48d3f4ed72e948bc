"""Python side of the bit-packed XNOR-popcount forward (``csrc/mnb_xnor.cu``) for wbwtab layers.

    pack_act     fp32 NCHW -> sign bit planes u32 [B][G][ceil(C/g / 32)][H][W]      (WB:11-36: sign(x), 0 -> +1)
    pack_weight  i16 levels {-1, 0, +1} -> sign / non-zero words + popcount / border tables   (WB:40-75, 98-146)
    conv         y = fmaf(popc(N) - 2 popc(N & (A ^ S)), alpha[k], bias[k])       (WB:181-195, forward only)

The integer sum is exact, so the result equals the packed-operand tensor-core forward bit for bit; which of the two runs a
given layer is decided from measurements (``harness/xnor_probe.py`` -> ``profiles/r2_xnor_vs_tc.md``)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L

_sup_cache = {}


def supported(sh):
    k = tuple(getattr(sh, f) for f, _ in sh._fields_)
    if k not in _sup_cache:
        _sup_cache[k] = L.load().mnb_xnor_supported(C.byref(sh)) == 1
    return _sup_cache[k]


def pack_act(x, groups):
    lib = L.load()
    b, c, h, w = x.shape
    nbytes = int(lib.mnb_xnor_act_bytes(b, c, h, w, groups))
    if nbytes < 0:
        raise ValueError("micronet_b200.xnor: channels not divisible by groups")
    out = torch.empty(nbytes // 4, dtype=torch.int32, device=x.device)
    L.check(lib.mnb_xnor_pack_act(x.data_ptr(), b, c, h, w, groups, out.data_ptr(), L.stream()), "xnor_pack_act")
    return out


def pack_weight(sh, w_int):
    lib = L.load()
    nbytes = int(lib.mnb_xnor_wimage_bytes(C.byref(sh)))
    if nbytes < 0:
        raise ValueError("micronet_b200.xnor: shape outside the cover of the XNOR-popcount convolution")
    img = torch.empty(nbytes // 4, dtype=torch.int32, device=w_int.device)
    L.check(lib.mnb_xnor_pack_weight(C.byref(sh), w_int.data_ptr(), img.data_ptr(), L.stream()), "xnor_pack_weight")
    return img


def conv(sh, a_bits, w_img, out, alpha=None, bias=None):
    return L.load().mnb_xnor_conv_fwd(C.byref(sh), a_bits.data_ptr(), w_img.data_ptr(), L.ptr(alpha), L.ptr(bias),
                                      out.data_ptr(), L.stream())
