"""GroupNorm+ELU forward / backward per layer shape of the B=4 192x640 step: cluster kernels vs two-pass kernels (CUDA events,
L2 flushed between runs).   python tools/gn_bench.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from packnet_sfm_b200 import functional as PF  # noqa: E402

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
shapes = [(4, 192, 640, 64), (4, 96, 320, 64), (4, 96, 320, 32), (4, 48, 160, 128), (4, 48, 160, 64), (4, 24, 80, 256), (4, 24, 80, 128),
          (4, 12, 40, 512), (4, 12, 40, 256), (4, 6, 20, 512), (4, 6, 20, 256)]


def timed(fn, iters=7):
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


modes = sys.argv[1:] or ["0", "1"]
print("%-22s %s" % ("shape", "  ".join("fwd/bwd us [PN_GN_CLUSTER=%s]" % m for m in modes)))
for B, H, W, C in shapes:
    x = (torch.rand(B, H, W, C, device=dev) - 0.3).requires_grad_(True)
    gm, bt = torch.rand(C, device=dev) + 0.5, torch.rand(C, device=dev) - 0.5
    gy = torch.rand(B, H, W, C, device=dev) - 0.5
    row = []
    for m in modes:
        os.environ["PN_GN_CLUSTER"] = m.split(":")[0]
        if ":" in m:
            os.environ["PN_GN_CLUSTER_MAX"] = m.split(":")[1]
        out = {}

        def fwd():
            out["y"] = PF.groupnorm_elu(x, gm, bt, 1e-5)

        def bwd():
            torch.autograd.grad(out["y"], x, gy, retain_graph=True)

        st = torch.cuda.Stream()
        with torch.cuda.stream(st):
            for _ in range(2):
                fwd(); bwd()
            torch.cuda.synchronize()
            gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()     # replayed graphs: no host enqueue gaps in the event time
            with torch.cuda.graph(gf):
                fwd()
            with torch.cuda.graph(gb):
                bwd()
            row.append("%7.1f /%7.1f" % (timed(gf.replay), timed(gb.replay)))
    print("%-22s %s" % ("%dx%dx%dx%d" % (B, H, W, C), "        ".join(row)))
