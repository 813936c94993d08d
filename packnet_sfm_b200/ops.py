"""Thin Python wrappers over the C-ABI layer ops (include/packnet_b200.h).  CUDA tensors only."""
import ctypes

import torch

from . import _lib
from ._lib_conv import ConvDesc, PRECISION_TF32X1, PRECISION_TF32X3, MODE_AUTO, MODE_PER_TAP, MODE_HALO  # noqa: F401


def tf32_residual(x):
    """lo = x - trunc_tf32(x) (same shape/strides as x)."""
    _lib.require_cuda(x)
    lo = torch.empty_like(x)
    _lib.check(_lib.lib().pn_tf32_residual(_lib.ptr(x), _lib.ptr(lo), x.numel(), _lib.current_stream()), "pn_tf32_residual")
    return lo


def pack_conv_weight(w_oihw, transposed=False, with_residual=False):
    """nn.Conv2d weight [Cout,Cin,k,k] -> packed K-major operand (and its tf32 residual)."""
    w = w_oihw.contiguous()
    _lib.require_cuda(w)
    cout, cin, k, _ = w.shape
    n = ctypes.c_size_t(0)
    lib = _lib.lib()
    _lib.check(lib.pn_conv2d_packed_weight_elems(cout, cin, k, int(transposed), ctypes.byref(n)), "pn_conv2d_packed_weight_elems")
    wp = torch.empty(int(n.value), dtype=torch.float32, device=w.device)
    lo = torch.empty_like(wp) if with_residual else None
    _lib.check(lib.pn_conv2d_pack_weight(_lib.ptr(w), _lib.ptr(wp), _lib.ptr(lo) if lo is not None else None, cout, cin, k,
                                         int(transposed), _lib.current_stream()), "pn_conv2d_pack_weight")
    return wp, lo


def conv2d_nhwc_packed(x, x_lo, wp, wp_lo, bias, cout, ksize, precision=PRECISION_TF32X3, mode=MODE_AUTO, debug_flags=0,
                       error_flag=None):
    """x: [B,H,W,Cin] contiguous NHWC storage; returns [B,H,W,Cout]."""
    _lib.require_cuda(x, wp)
    B, H, W, Cin = x.shape
    y = torch.empty(B, H, W, cout, dtype=torch.float32, device=x.device)
    d = ConvDesc(B, H, W, Cin, cout, ksize, precision, mode, debug_flags)
    _lib.check(_lib.lib().pn_conv2d_forward(
        ctypes.byref(d), _lib.ptr(x), _lib.ptr(x_lo) if x_lo is not None else None, _lib.ptr(wp),
        _lib.ptr(wp_lo) if wp_lo is not None else None, _lib.ptr(bias) if bias is not None else None, _lib.ptr(y),
        _lib.ptr(error_flag) if error_flag is not None else None, _lib.current_stream()), "pn_conv2d_forward")
    return y


def conv2d_nhwc(x, w_oihw, bias=None, precision=PRECISION_TF32X3, mode=MODE_AUTO, debug_flags=0, error_flag=None):
    """Convenience (tests): packs the weight on every call."""
    x = x.contiguous()
    three = precision == PRECISION_TF32X3
    wp, wlo = pack_conv_weight(w_oihw, False, three)
    xlo = tf32_residual(x) if three else None
    return conv2d_nhwc_packed(x, xlo, wp, wlo, bias, w_oihw.shape[0], w_oihw.shape[2], precision, mode, debug_flags, error_flag)
