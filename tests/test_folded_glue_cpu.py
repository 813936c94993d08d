"""The REAL Python glue of the folded pack block -- ctypes descriptors, pointer offsets into the border lines, the nine
fold calls with their layouts and accumulate flags, the in-place frame function and both autograd backward chains --
driven end to end on the CPU with tests/kernel_mirrors.MirrorLib standing in for libpacknet_b200.so (the kernels
themselves need the GPU tier).  What is checked: values and every gradient against the reference composition
packing -> Conv3d -> pad -> Conv2d (layers01.py:239-247)."""
import pytest
import torch
import torch.nn.functional as F

from kernel_mirrors import MirrorLib
from oracle import packnet_oracle as PO
from packnet_sfm_b200 import _lib, folded


def _conv_nhwc(x, w, b):
    return F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=w.shape[-1] // 2).permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("case", [(1, 2, 8, 12, 3, 3), (2, 1, 10, 12, 2, 5)])
def test_kernel_path_glue_with_mirror_library(case, monkeypatch):
    B, C, H, W, Co, k = case
    lib = MirrorLib()
    monkeypatch.setattr(_lib, "lib", lambda: lib)
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    monkeypatch.setattr(_lib, "current_stream", lambda: None)
    monkeypatch.setattr(folded, "_use_kernels", lambda t: True)
    g = torch.Generator().manual_seed(B + C + k)
    x = (torch.rand(B, H, W, C, generator=g) - 0.5).requires_grad_(True)
    w2 = ((torch.rand(Co, 32 * C, k, k, generator=g) - 0.5) * 0.3).requires_grad_(True)
    b2 = (torch.rand(Co, generator=g) - 0.5).requires_grad_(True)
    w3 = (torch.rand(8, 1, 3, 3, 3, generator=g) - 0.5).requires_grad_(True)
    b3 = (torch.rand(8, generator=g) - 0.5).requires_grad_(True)
    z = folded.pack_conv_folded(x, w2, b2, w3, b3, _conv_nhwc)
    gz = torch.rand(z.shape, generator=g) - 0.5
    z.backward(gz)
    assert lib.calls.count("fold_fwd") == 9 and lib.calls.count("fold_bwd") == 9
    assert lib.calls.count("frame_fwd") == 1 and lib.calls.count("frame_bwd") == 1
    xd, w2d, b2d, w3d, b3d = (t.detach().double().requires_grad_(True) for t in (x, w2, b2, w3, b3))
    t = PO.conv3d_features(PO.packing(xd.permute(0, 3, 1, 2)), w3d, b3d)
    zr = F.conv2d(F.pad(t, [k // 2] * 4), w2d, b2d).permute(0, 2, 3, 1)
    zr.backward(gz.double())
    for name, a, b in (("z", z, zr), ("gx", x.grad, xd.grad), ("gw2", w2.grad, w2d.grad), ("gb2", b2.grad, b2d.grad),
                       ("gw3", w3.grad, w3d.grad), ("gb3", b3.grad, b3d.grad)):
        err = float((a.detach().double() - b.detach()).abs().max() / b.detach().abs().max())
        assert err < 2e-5, (name, err)
