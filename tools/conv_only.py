"""One convolution through the engine, three times (target of `ncu -k regex:conv_`).
    python tools/conv_only.py [B H W Cin Cout k [fwd|dgrad|wgrad] [debug_flags]]     default: the folded pack1 forward"""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from packnet_sfm_b200 import _lib, functional as PF  # noqa: E402
from packnet_sfm_b200._lib_conv import ConvDesc  # noqa: E402

a = sys.argv[1:]
B, H, W, Cin, Cout, k = (int(v) for v in a[:6]) if len(a) >= 6 else (4, 96, 320, 256, 64, 7)
what = a[6] if len(a) > 6 else "fwd"
dbg = int(a[7]) if len(a) > 7 else 0
prec = PF.PRECISION_BF16X3
dev = torch.device("cuda:0")
lib = _lib.lib()
x = torch.rand(B, H, W, Cin, device=dev) - 0.5
g = torch.rand(B, H, W, Cout, device=dev) - 0.5
w = (torch.rand(Cout, Cin, k, k, device=dev) - 0.5) * 0.01
wp, wlo = PF._pack_weight(w, False, prec)
xh, xlo = PF._operands(x, prec)
gh, glo = PF._operands(g, prec)
d = ConvDesc(B, H, W, Cin, Cout, k, prec, 0, dbg)
n = ctypes.c_size_t(0)
lib.pn_conv2d_wgrad_packed_elems(Cout, Cin, k, prec, ctypes.byref(n))
dwp = torch.empty(int(n.value), device=dev)
for _ in range(3):
    if what == "fwd":
        y = torch.empty(B, H, W, Cout, device=dev)
        _lib.check(lib.pn_conv2d_forward(ctypes.byref(d), _lib.ptr(xh), _lib.ptr(xlo), _lib.ptr(wp), _lib.ptr(wlo), None, _lib.ptr(y),
                                         _lib.ptr(PF.error_flag()), _lib.current_stream()), "fwd")
    elif what == "dgrad":
        y = PF._conv_dgrad(gh, glo, wp, wlo, B, H, W, Cin, Cout, k, prec)
    else:
        _lib.check(lib.pn_conv2d_wgrad(ctypes.byref(d), _lib.ptr(xh), _lib.ptr(xlo), _lib.ptr(gh), _lib.ptr(glo), _lib.ptr(dwp),
                                       _lib.ptr(PF.error_flag()), _lib.current_stream()), "wgrad")
        y = dwp
torch.cuda.synchronize()
print("ok", what, float(y.abs().mean()))
