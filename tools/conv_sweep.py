"""Round-2 tuning sweep of the convolution engine (run under gpurun): for the low-resolution layers of the step, the time
and the error (against float64 F.conv2d) of the forward kernel with the tile height forced to 8 / 4 / 2 rows
(debug_flags bits 16..19; 4 rows x 2 images fills every MMA row of a 12-row map, at the price of per-tap staging), and the
shapes of the folded pack layers (packnet_sfm_b200/folded.py) next to the layers they replace."""
import ctypes
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from packnet_sfm_b200 import _lib, functional as PF  # noqa: E402
from packnet_sfm_b200._lib_conv import ConvDesc  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.lib()
P = PF.PRECISION_BF16X3


def run_case(tag, B, H, W, Cin, Cout, k, dbg=0, check=True):
    torch.manual_seed(1)
    x = torch.rand(B, H, W, Cin, device=dev) - 0.5
    w = (torch.rand(Cout, Cin, k, k, device=dev) - 0.5) * (2.0 / (Cin * k * k) ** 0.5)
    wp, wlo = PF._pack_weight(w, False, P)
    xh, xl = PF._operands(x, P)
    y = torch.empty(B, H, W, Cout, device=dev)
    d = ConvDesc(B, H, W, Cin, Cout, k, P, 0, dbg)

    def run():
        rc = lib.pn_conv2d_forward(ctypes.byref(d), _lib.ptr(xh), PF._p(xl), _lib.ptr(wp), PF._p(wlo), None, _lib.ptr(y),
                                   _lib.ptr(PF.error_flag()), _lib.current_stream())
        if rc:
            raise RuntimeError(lib.pn_last_error_string().decode())
    try:
        for _ in range(2):
            run()
        torch.cuda.synchronize()
    except RuntimeError as e:
        print("%-58s refused: %s" % (tag, str(e)[:80]), flush=True)
        return
    ts = []
    for _ in range(7):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    err = float("nan")
    if check:
        yr = F.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), padding=k // 2).permute(0, 2, 3, 1)
        err = float((y.double() - yr).norm() / yr.norm())
    print("%-58s %8.4f ms %7.1f TFLOP/s  rel err %.2e" % (tag, ms, 2.0 * B * H * W * Cout * Cin * k * k / ms / 1e9, err), flush=True)


B = 4
print("== tile height sweep on the low-resolution layers (default picks 8 rows unless the map has <= 4)")
for (H, W, ci, co, k) in ((12, 40, 512, 512, 3), (12, 40, 768, 512, 3), (24, 80, 256, 256, 3), (6, 20, 512, 256, 3),
                          (12, 40, 512, 512, 1), (6, 20, 16384, 512, 3), (12, 40, 8192, 256, 3)):
    for th in (0, 8, 4, 2):
        run_case("%dx%d %d->%d k%d  tile rows %s" % (H, W, ci, co, k, th or "default"), B, H, W, ci, co, k, dbg=th << 16,
                 check=ci <= 1024)
print("== folded pack layers (forward and data-gradient shapes) next to the layers they replace")
for name, (H, W, n, co, k) in (("pack1", (96, 320, 256, 64, 5)), ("pack2", (48, 160, 256, 64, 3)), ("pack3", (24, 80, 512, 128, 3)),
                               ("pack4", (12, 40, 1024, 256, 3)), ("pack5", (6, 20, 2048, 512, 3))):
    run_case("%s today    %d->%d k%d" % (name, 8 * n, co, k), B, H, W, 8 * n, co, k, check=False)
    run_case("%s folded   %d->%d k%d" % (name, n, co, k + 2), B, H, W, n, co, k + 2)
    run_case("%s today    dgrad %d->%d k%d" % (name, co, 8 * n, k), B, H, W, co, 8 * n, k, check=False)
    run_case("%s folded   dgrad %d->%d k%d" % (name, co, n, k + 2), B, H, W, co, n, k + 2)
