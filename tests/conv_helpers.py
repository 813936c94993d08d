"""Test / probe helper: one-call convolution through the C-ABI (packs the weight and splits the activation every call)."""
import ctypes

import torch

from packnet_sfm_b200 import _lib
from packnet_sfm_b200._lib_conv import (ConvDesc, PRECISION_TF32X1, PRECISION_BF16X1, PRECISION_TF32X3, PRECISION_BF16X3,  # noqa: F401
                        MODE_AUTO, MODE_PER_TAP, MODE_HALO)


def conv2d_nhwc(x, w_oihw, bias=None, precision=None, mode=MODE_AUTO, debug_flags=0, error_flag=None):
    """Convenience (tests / probes): packs the weight and splits the activation on every call."""
    from packnet_sfm_b200 import functional as PF
    precision = PF.get_precision() if precision is None else precision
    x = x.contiguous()
    wp, wlo = PF._pack_weight(PF._pad_channels(w_oihw.contiguous(), x.shape[3]), False, precision)
    xh, xl = PF._operands(x, precision)
    B, H, W, Cin = x.shape
    y = torch.empty(B, H, W, w_oihw.shape[0], dtype=torch.float32, device=x.device)
    d = ConvDesc(B, H, W, Cin, w_oihw.shape[0], w_oihw.shape[2], precision, mode, debug_flags)
    flag = error_flag if error_flag is not None else PF.error_flag()
    _lib.check(_lib.lib().pn_conv2d_forward(ctypes.byref(d), _lib.ptr(xh), _lib.ptr(xl) if xl is not None else None,
                                            _lib.ptr(wp), _lib.ptr(wlo) if wlo is not None else None,
                                            _lib.ptr(bias) if bias is not None else None, _lib.ptr(y), _lib.ptr(flag),
                                            _lib.current_stream()), "pn_conv2d_forward")
    return y
