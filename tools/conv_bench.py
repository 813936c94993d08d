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
for gsz in (1, 2, 3):
    bench("pack1 bf16x3 halo  group %d" % gsz, 4, 96, 320, 2048, 64, 5, B3, 2, dbg=gsz << 8)
for gsz in (1, 2, 4, 6):
    bench("pack1 bf16x1 halo  group %d" % gsz, 4, 96, 320, 2048, 64, 5, B1, 2, dbg=gsz << 8)
for gsz in (1, 2, 3):
    bench("k3 256->256 24x80 bf16x3 group %d" % gsz, 4, 24, 80, 256, 256, 3, B3, 0, dbg=gsz << 8)
bench("pack5 16384->512 6x20 bf16x3", 4, 6, 20, 16384, 512, 3, B3, 0)
bench("512->512 12x40 bf16x3", 4, 12, 40, 512, 512, 3, B3, 0)
bench("pack1 bf16x3 per-tap", 4, 96, 320, 2048, 64, 5, B3, 1)
bench("pack1 bf16x1 halo", 4, 96, 320, 2048, 64, 5, B1, 2)
bench("pack1 tf32x1 halo", 4, 96, 320, 2048, 64, 5, T1, 2)
bench("pack1-like Cin512 bf16x3 halo", 4, 96, 320, 512, 64, 5, B3, 2)
bench("pack1-like k3 bf16x3 halo", 4, 96, 320, 2048, 64, 3, B3, 2)
bench("pack1-like k1 bf16x3", 4, 96, 320, 2048, 64, 1, B3, 1)
bench("pack1-like Cout128 bf16x3 halo", 4, 96, 320, 2048, 128, 5, B3, 2)
bench("pack1-like B1 (240 tiles) bf16x3 halo", 1, 96, 320, 2048, 64, 5, B3, 2)
bench("148 tiles exactly: bf16x3 halo", 1, 16 * 37, 32, 2048, 64, 5, B3, 2)
bench("conv1-like 64->64 k7 bf16x3", 4, 192, 640, 64, 64, 7, B3, 2)
bench("res 128->128 k3 48x160 bf16x3", 4, 48, 160, 128, 128, 3, B3, 2)
