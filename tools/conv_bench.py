"""Timing experiments on the pack1-shaped convolution (which stage of the pipeline limits it?)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from packnet_sfm_b200 import _lib, functional as PF  # noqa: E402
from packnet_sfm_b200._lib_conv import ConvDesc  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.lib()


def bench(tag, B, H, W, Cin, Cout, k, prec, mode, iters=5, dbg=0):
    x = torch.rand(B, H, W, Cin, device=dev) - 0.5
    w = (torch.rand(Cout, Cin, k, k, device=dev) - 0.5) * 0.01
    wp, wlo = PF._pack_weight(w, False, prec)
    xh, xl = PF._operands(x, prec)
    y = torch.empty(B, H, W, Cout, device=dev)
    d = ConvDesc(B, H, W, Cin, Cout, k, prec, mode, dbg)

    def run():
        _lib.check(lib.pn_conv2d_forward(ctypes.byref(d), _lib.ptr(xh), PF._p(xl), _lib.ptr(wp), PF._p(wlo), None, _lib.ptr(y),
                                         _lib.ptr(PF.error_flag()), _lib.current_stream()), "conv")
    for _ in range(2):
        run()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    ms = ts[len(ts) // 2]
    flops = 2.0 * B * H * W * Cout * Cin * k * k
    kc = 64 if PF.is_bf16(prec) else 32
    tiles = B * ((H + 15) // 16) * ((W + 7) // 8) * ((Cout + 127) // 128)
    stages = ((Cin + kc - 1) // kc) * k * k
    waves = -(-tiles // 148)
    clk_per_stage = ms * 1e-3 * 1.965e9 / (waves * stages)
    print("%-44s %8.3f ms  %7.1f TFLOP/s  tiles %5d waves %3d stages/CTA %5d  ~%6.0f clk/stage" % (
        tag, ms, flops / ms / 1e9, tiles, waves, stages, clk_per_stage), flush=True)


B3, B1, T3, T1 = PF.PRECISION_BF16X3, PF.PRECISION_BF16X1, PF.PRECISION_TF32X3, PF.PRECISION_TF32X1
bench("pack1 bf16x3 halo", 4, 96, 320, 2048, 64, 5, B3, 2)
for tag, dbg in (("persistent", 0), ("one tile per CTA", 4096)):
    bench("pack1 bf16x3 halo  %s" % tag, 4, 96, 320, 2048, 64, 5, B3, 2, dbg=dbg)
    bench("64->64 k3 96x320 bf16x3  %s" % tag, 4, 96, 320, 64, 64, 3, B3, 0, dbg=dbg)
    bench("64->64 k7 192x640 bf16x3  %s" % tag, 4, 192, 640, 64, 64, 7, B3, 0, dbg=dbg)
    bench("64->136 k3 192x640 bf16x3  %s" % tag, 4, 192, 640, 64, 136, 3, B3, 0, dbg=dbg)
    bench("256->256 k3 24x80 bf16x3  %s" % tag, 4, 24, 80, 256, 256, 3, B3, 0, dbg=dbg)
    bench("64->2048 k5 96x320 bf16x3 (pack1 dgrad)  %s" % tag, 4, 96, 320, 64, 2048, 5, B3, 0, dbg=dbg)
