"""Decode what the MN-major wgrad MMA computes on structured inputs (bring-up)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from packnet_sfm_b200 import _lib, functional as PF  # noqa: E402
from packnet_sfm_b200._lib_conv import ConvDesc  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.lib()


def run(tag, B, H, W, Cin, Cout, k, debug, x, g):
    n = ctypes.c_size_t(0)
    lib.pn_conv2d_packed_weight_elems(Cout, Cin, k, 0, ctypes.byref(n))
    dwp = torch.full((int(n.value),), 7.0, device=dev)
    d = ConvDesc(B, H, W, Cin, Cout, k, 1, 0, debug)
    rc = lib.pn_conv2d_wgrad(ctypes.byref(d), _lib.ptr(x), None, _lib.ptr(g), None, _lib.ptr(dwp), _lib.ptr(PF.error_flag()),
                             _lib.current_stream())
    torch.cuda.synchronize()
    kp = (Cin + 31) // 32 * 32
    D = dwp.view(Cout, k * k, kp)
    ref = torch.einsum("bhwi,bhwo->oi", x.double(), g.double()) if k == 1 else None
    print("== %s rc=%d flag=%x  |D|max=%.4g nonzero=%d/%d" % (tag, rc, PF.read_error_flag(), float(D.abs().max()),
                                                         int((D != 0).sum()), D.numel()))
    if ref is not None:
        got = D[:, 0, :Cin].double()
        print("   rel err vs einsum: %.3e ; vs transposed: %.3e" % (float((got - ref).norm() / ref.norm()),
                                                                  float((got - ref.t()).norm() / ref.norm()) if Cin == Cout else -1))
        print("   got[0:4,0:6]=\n%s\n   ref[0:4,0:6]=\n%s" % (got[:4, :6], ref[:4, :6]))


B, H, W, C = 1, 8, 8, 32
x = torch.zeros(B, H, W, C, device=dev)
g = torch.zeros(B, H, W, C, device=dev)
p = torch.arange(H * W, device=dev)
x.view(-1, C)[p, p % C] = 1.0                         # one-hot over channels
g.view(-1, C)[:] = (p[:, None] * 100 + torch.arange(C, device=dev)[None, :]).float()
for dbg in (0, 4):
    run("onehot k1 debug=%d" % dbg, B, H, W, C, C, 1, dbg, x, g)
torch.manual_seed(0)
x = torch.rand(B, H, W, C, device=dev)
g = torch.rand(B, H, W, C, device=dev)
for dbg in (0, 4):
    run("random k1 debug=%d" % dbg, B, H, W, C, C, 1, dbg, x, g)
x = torch.rand(1, 16, 16, 64, device=dev)
g = torch.rand(1, 16, 16, 64, device=dev)
run("random 16x16 C64 k1", 1, 16, 16, 64, 64, 1, 0, x, g)
x = torch.rand(2, 20, 24, 96, device=dev)
g = torch.rand(2, 20, 24, 64, device=dev)
run("random B2 20x24 Cin96 Cout64 k1", 2, 20, 24, 96, 64, 1, 0, x, g)
