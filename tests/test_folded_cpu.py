"""Folded pack block (packnet_sfm_b200/folded.py): the algebra -- one (k+2)x(k+2) convolution of the space-to-depth
tensor plus exact frame terms -- against the reference composition packing -> Conv3d -> pad -> Conv2d
(layers01.py:239-247, restated in oracle/packnet_oracle.py) in float64 on the CPU: values and every gradient."""
import pytest
import torch
import torch.nn.functional as F

from oracle import packnet_oracle as PO
from packnet_sfm_b200 import folded


def _conv_nhwc(x, w, b):
    """stand-in for functional.conv2d (the tcgen05 engine) on the CPU: NHWC in, NHWC out, zero pad k//2"""
    return F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=w.shape[-1] // 2).permute(0, 2, 3, 1).contiguous()


def _reference(x_nchw, w2, b2, w3, b3, k):
    t = PO.conv3d_features(PO.packing(x_nchw), w3, b3)
    return F.conv2d(F.pad(t, [k // 2] * 4), w2, b2)


CASES = [
    # B, C, H, W, Co, k
    (2, 4, 12, 16, 8, 3),
    (1, 8, 14, 10, 16, 5),
    (2, 2, 6, 6, 4, 3),        # packed map 3x3 = exactly the frame of a 3x3 kernel
    (1, 4, 10, 12, 4, 5),      # packed map 5x6: the minimum for k = 5
    (1, 16, 24, 40, 16, 5),    # pack1-like aspect
    (3, 6, 8, 20, 12, 3),
]


@pytest.mark.parametrize("case", CASES)
def test_folded_pack_conv_equals_reference_composition(case):
    B, C, H, W, Co, k = case
    g = torch.Generator().manual_seed(B * 100 + C + H + k)
    dd = dict(dtype=torch.float64, generator=g)
    x = (torch.rand(B, H, W, C, **dd) - 0.5).requires_grad_(True)
    w2 = ((torch.rand(Co, 32 * C, k, k, **dd) - 0.5) * 0.2).requires_grad_(True)
    b2 = (torch.rand(Co, **dd) - 0.5).requires_grad_(True)
    w3 = (torch.rand(8, 1, 3, 3, 3, **dd) - 0.5).requires_grad_(True)
    b3 = (torch.rand(8, **dd) - 0.5).requires_grad_(True)
    gz = torch.rand(B, H // 2, W // 2, Co, **dd) - 0.5

    z = folded.pack_conv_folded(x, w2, b2, w3, b3, _conv_nhwc)
    z.backward(gz)
    got = [z.detach()] + [t.grad.clone() for t in (x, w2, b2, w3, b3)]

    xr, w2r, b2r, w3r, b3r = (t.detach().clone().requires_grad_(True) for t in (x, w2, b2, w3, b3))
    zr = _reference(xr.permute(0, 3, 1, 2), w2r, b2r, w3r, b3r, k).permute(0, 2, 3, 1)
    zr.backward(gz)
    want = [zr.detach()] + [t.grad for t in (xr, w2r, b2r, w3r, b3r)]

    for name, a, b in zip(("z", "gx", "gw2", "gb2", "gw3", "gb3"), got, want):
        err = float((a - b).abs().max() / b.abs().max())
        assert err < 1e-11, (name, err)


def test_folded_pack_conv_fp32_error_is_rounding_only():
    """fp32 run of the folded form vs the float64 reference: the fold changes the summation order, not the bar."""
    torch.manual_seed(3)
    B, C, H, W, Co, k = 1, 16, 24, 32, 16, 5
    x = torch.rand(B, H, W, C) - 0.5
    w2 = (torch.rand(Co, 32 * C, k, k) - 0.5) * (2.0 / (32 * C * k * k) ** 0.5)
    b2, w3, b3 = torch.rand(Co) - 0.5, torch.rand(8, 1, 3, 3, 3) - 0.5, torch.rand(8) - 0.5
    z = folded.pack_conv_folded(x, w2, b2, w3, b3, _conv_nhwc)
    zr = _reference(x.double().permute(0, 3, 1, 2), w2.double(), b2.double(), w3.double(), b3.double(), k).permute(0, 2, 3, 1)
    assert float((z.double() - zr).abs().max() / zr.abs().max()) < 2e-6


def test_folded_rejects_maps_smaller_than_the_frame():
    x = torch.zeros(1, 4, 8, 2)
    with pytest.raises(ValueError):
        folded.pack_conv_folded(x, torch.zeros(4, 64, 5, 5), torch.zeros(4), torch.zeros(8, 1, 3, 3, 3), torch.zeros(8), _conv_nhwc)
