"""Randomised parity sweep of the folded pack block's kernel path (fold + frame kernels under the host emulation, PyTorch
convolution for the O(area) part) against the reference composition packing -> Conv3d -> pad -> Conv2d in float64: random
batch, channels, map sizes down to the frame itself, output channels, k in {3, 5}; values and all five gradients.
TEST INFRASTRUCTURE; not collected by pytest:  python tests/emu/fuzz_folded.py <seed> <seconds>"""
import ctypes, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.nn.functional as F
from packnet_sfm_b200 import _lib, _lib_conv, folded
from oracle import packnet_oracle as PO
lib = ctypes.CDLL(os.path.join(ROOT, 'tests', 'emu', '_build', 'libpacknet_emu.so'))
_lib._declare(lib); _lib_conv.declare(lib)
_lib.lib = lambda: lib; _lib.require_cuda = lambda *a: None; _lib.current_stream = lambda: None
folded._use_kernels = lambda t: True


def conv(x, w, b):
    return F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=w.shape[-1] // 2).permute(0, 2, 3, 1).contiguous()


random.seed(int(sys.argv[1])); t_end = time.time() + float(sys.argv[2]); n = 0
while time.time() < t_end:
    k = random.choice([3, 5]); m = k // 2
    B = random.choice([1, 1, 2, 3]); C = random.choice([1, 2, 3, 5, 8]); Co = random.choice([1, 2, 3, 5])
    h = random.randint(2 * m + 1, 2 * m + 6); w = random.randint(2 * m + 1, 2 * m + 9)
    g = torch.Generator().manual_seed(random.randint(0, 1 << 30))
    x = (torch.rand(B, 2 * h, 2 * w, C, generator=g) - 0.5).requires_grad_(True)
    w2 = ((torch.rand(Co, 32 * C, k, k, generator=g) - 0.5) * 0.2).requires_grad_(True)
    b2 = (torch.rand(Co, generator=g) - 0.5).requires_grad_(True)
    w3 = (torch.rand(8, 1, 3, 3, 3, generator=g) - 0.5).requires_grad_(True)
    b3 = (torch.rand(8, generator=g) - 0.5).requires_grad_(True)
    z = folded.pack_conv_folded(x, w2, b2, w3, b3, conv)
    gz = torch.rand(z.shape, generator=g) - 0.5
    z.backward(gz)
    xd, w2d, b2d, w3d, b3d = (t.detach().double().requires_grad_(True) for t in (x, w2, b2, w3, b3))
    t = PO.conv3d_features(PO.packing(xd.permute(0, 3, 1, 2)), w3d, b3d)
    zr = F.conv2d(F.pad(t, [m] * 4), w2d, b2d).permute(0, 2, 3, 1)
    zr.backward(gz.double())
    errs = [float((a.detach().double() - b.detach()).abs().max() / b.detach().abs().max())
            for a, b in ((z, zr), (x.grad, xd.grad), (w2.grad, w2d.grad), (b2.grad, b2d.grad), (w3.grad, w3d.grad), (b3.grad, b3d.grad))]
    n += 1
    if max(errs) > 5e-5:
        print("MISMATCH", (B, C, h, w, Co, k), errs, flush=True)
print("cases", n)
