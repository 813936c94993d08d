"""Bring-up probe for the tcgen05 implicit-GEMM convolution: prints error statistics per configuration
against torch fp32 (cudnn.allow_tf32=False) and against a tf32-truncated-operand emulation.
Usage: python tools/conv_probe.py [group]   (groups: pertap, halo, x3, all)"""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import conv_helpers as ops  # noqa: E402


def trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def run(tag, B, H, W, Cin, Cout, k, precision, mode, debug=0, seed=0):
    torch.manual_seed(seed)
    dev = torch.device("cuda:0")
    x = (torch.rand(B, H, W, Cin, device=dev) - 0.5)
    w = (torch.rand(Cout, Cin, k, k, device=dev) - 0.5) * (2.0 / (Cin * k * k) ** 0.5)
    bias = torch.rand(Cout, device=dev) - 0.5
    flag = torch.zeros(4, dtype=torch.int32, device=dev)
    try:
        y = ops.conv2d_nhwc(x, w, bias, precision, mode, debug, flag)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("%-34s EXCEPTION %s flag=%s" % (tag, str(e)[:150], hex(int(flag[0].item()) & 0xFFFFFFFF) if False else "?"))
        return False
    xn = x.permute(0, 3, 1, 2)
    ref = F.conv2d(xn.double(), w.double(), bias.double(), padding=k // 2).permute(0, 2, 3, 1)
    emu = F.conv2d(trunc(x).permute(0, 3, 1, 2).double(), trunc(w).double(), bias.double(), padding=k // 2).permute(0, 2, 3, 1)
    e_ref = float((y.double() - ref).norm() / ref.norm())
    e_emu = float((y.double() - emu).norm() / emu.norm())
    mx = float((y.double() - ref).abs().max())
    # where do the errors sit (helps decode layout bugs)
    bad = ((y.double() - (emu if precision == 1 else ref)).abs() > 1e-3)
    nbad = int(bad.sum())
    info = ""
    if nbad:
        idx = bad.nonzero()[:4].tolist()
        info = " bad=%d/%d first=%s" % (nbad, bad.numel(), idx)
    print("%-34s rel_l2(fp32)=%.3e rel_l2(tf32emu)=%.3e maxabs=%.3e flag=%x%s" % (
        tag, e_ref, e_emu, mx, int(flag[0].item()) & 0xFFFFFFFF, info))
    return True


def run_bwd(tag, B, H, W, Cin, Cout, k, precision, which, seed=0):
    """which: 'dgrad' or 'wgrad' -- exercises the backward kernels in isolation."""
    from packnet_sfm_b200 import functional as PF
    torch.manual_seed(seed)
    dev = torch.device("cuda:0")
    PF.set_precision(precision)
    x = (torch.rand(B, H, W, Cin, device=dev) - 0.5).requires_grad_(which == "dgrad")
    w = ((torch.rand(Cout, Cin, k, k, device=dev) - 0.5) * (2.0 / (Cin * k * k) ** 0.5)).requires_grad_(which == "wgrad")
    gy = torch.rand(B, H, W, Cout, device=dev) - 0.5
    try:
        y = PF.conv2d(x, w, None)
        torch.cuda.synchronize()
        y.backward(gy)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        print("%-40s EXCEPTION %s  pipeline-timeout flag=%x" % (tag, str(e)[:120], PF.read_error_flag()))
        return False
    xd, wd = x.detach().double().requires_grad_(True), w.detach().double().requires_grad_(True)
    yr = F.conv2d(xd.permute(0, 3, 1, 2), wd, None, padding=k // 2).permute(0, 2, 3, 1)
    yr.backward(gy.double())
    got, ref = (x.grad, xd.grad) if which == "dgrad" else (w.grad, wd.grad)
    print("%-40s %s rel_l2=%.3e" % (tag, which, float((got.double() - ref).norm() / ref.norm())))
    return True


def main():
    group = sys.argv[1] if len(sys.argv) > 1 else "all"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    X1, X3 = ops.PRECISION_TF32X1, ops.PRECISION_TF32X3
    PT, HL = ops.MODE_PER_TAP, ops.MODE_HALO
    if group in ("pertap", "all"):
        run("gemm 128x16x32 k1", 1, 16, 8, 32, 16, 1, X1, PT)
        run("gemm k1 Cin64 Cout64 2x2 tiles", 1, 32, 16, 64, 64, 1, X1, PT)
        run("k3 pertap 1 tile", 1, 16, 8, 32, 32, 3, X1, PT)
        run("k3 pertap ragged B2 40x36", 2, 40, 36, 64, 48, 3, X1, PT)
        run("k5 pertap Cin96 Cout64", 1, 32, 24, 96, 64, 5, X1, PT)
        run("k3 pertap fold H6 W20 B4", 4, 6, 20, 64, 128, 3, X1, PT)
        run("k3 pertap Cin36 (ragged K)", 1, 16, 16, 36, 32, 3, X1, PT)
        run("k1 Cout256", 1, 16, 16, 64, 256, 1, X1, PT)
        run("k3 Cout512 (2 n-tiles)", 1, 16, 16, 64, 512, 3, X1, PT)
    if group in ("halo", "all"):
        run("k3 halo 1 tile", 1, 16, 8, 32, 32, 3, X1, HL)
        run("k3 halo 1 tile boff=addr", 1, 16, 8, 32, 32, 3, X1, HL, 1)
        run("k5 halo B2 48x40 Cin64", 2, 48, 40, 64, 64, 5, X1, HL)
        run("k5 halo boff=addr", 2, 48, 40, 64, 64, 5, X1, HL, 1)
        run("k7 halo 32x24", 1, 32, 24, 64, 64, 7, X1, HL)
    if group in ("bf16", "all"):
        B3 = ops.PRECISION_BF16X3
        run("bf16x3 gemm k1", 1, 16, 8, 64, 16, 1, B3, PT)
        run("bf16x3 k3 pertap", 2, 40, 36, 64, 48, 3, B3, PT)
        run("bf16x3 k3 halo", 2, 40, 36, 64, 48, 3, B3, HL)
        run("bf16x3 k5 halo", 2, 48, 40, 64, 64, 5, B3, HL)
        run("bf16x3 k3 pertap K=2048", 1, 16, 16, 2048, 64, 3, B3, PT)
        run("bf16x3 fold H6 W20 B4", 4, 6, 20, 64, 128, 3, B3, PT)
        run("bf16x3 Cin40 ragged", 1, 16, 16, 40, 32, 3, B3, HL)
        run("bf16x1 k3", 1, 16, 16, 64, 32, 3, ops.PRECISION_BF16X1, HL)
    if group in ("x3", "all"):
        run("x3 gemm k1", 1, 16, 8, 32, 16, 1, X3, PT)
        run("x3 k3 pertap", 2, 40, 36, 64, 48, 3, X3, PT)
        run("x3 k5 halo", 2, 48, 40, 64, 64, 5, X3, HL)
        run("x3 k3 pertap K=2048", 1, 16, 16, 2048, 64, 3, X3, PT)
        run("x1 k3 pertap K=2048", 1, 16, 16, 2048, 64, 3, X1, PT)


def main_bwd(group):
    torch.backends.cudnn.allow_tf32 = False
    X1, X3 = ops.PRECISION_TF32X1, ops.PRECISION_BF16X3
    cases = [(2, 40, 36, 64, 48, 3), (1, 32, 24, 96, 64, 5), (4, 6, 20, 64, 128, 3), (1, 16, 16, 40, 32, 3),
             (1, 16, 16, 64, 512, 3), (1, 32, 24, 64, 64, 7), (2, 12, 40, 256, 256, 3)]
    only = int(sys.argv[2]) if len(sys.argv) > 2 else -1
    for i, c in enumerate(cases):
        if only >= 0 and i != only:
            continue
        for prec in (X1, X3):
            run_bwd("%s x%d %s" % (group, prec, c), *c, prec, group)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] in ("dgrad", "wgrad"):
        main_bwd(sys.argv[1])
        sys.exit(0)
    main()
