"""GPU tier (experimental, PN_EXPERIMENTAL=1): the staged first-layer re-composition -- functional.conv2d_im2col on the
tensor-core engine against the direct convolution in float64, and PackNet01 with it against the reference's golden depth
maps (bar 1e-3)."""
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle import packnet_oracle as PO

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(os.environ.get("PN_EXPERIMENTAL") != "1", reason="staged re-compositions: set PN_EXPERIMENTAL=1")]
DEV = "cuda:0"


@pytest.mark.parametrize("shape", [(2, 32, 64, 3, 64, 5), (4, 192, 640, 3, 64, 5), (1, 16, 24, 3, 32, 3)])
def test_im2col_first_layer_on_the_engine(shape):
    from packnet_sfm_b200 import functional as PF
    B, H, W, cin, cout, k = shape
    g = torch.Generator().manual_seed(H + k)
    x = torch.rand(B, H, W, cin, generator=g)
    w = (torch.rand(cout, cin, k, k, generator=g) - 0.5) * 0.2
    b = torch.rand(cout, generator=g) - 0.5
    wd, bd = w.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    y = PF.conv2d_im2col(x.to(DEV), wd, bd)
    gy = torch.rand(y.shape, generator=g) - 0.5
    y.backward(gy.to(DEV))
    torch.cuda.synchronize()
    wr, br = w.double().requires_grad_(True), b.double().requires_grad_(True)
    yr = F.conv2d(x.double().permute(0, 3, 1, 2), wr, br, padding=k // 2).permute(0, 2, 3, 1)
    yr.backward(gy.double())
    assert rel_l2(y.detach().cpu(), yr.detach()) < 2e-5
    assert rel_l2(wd.grad.cpu(), wr.grad) < 1e-4 and rel_l2(bd.grad.cpu(), br.grad) < 1e-4


def test_packnet01_with_im2col_first_layer_matches_reference_golden():
    from packnet_sfm_b200 import functional as PF
    from packnet_sfm_b200.networks import PackNet01
    z = load_golden("packnet01_64x96")
    net = PackNet01(version="1A")
    net.load_state_dict(PO.packnet01_state_dict(seed=42, randomize_affine=True), strict=True)
    net = net.to(DEV).train()
    prev = PF.set_im2col_first(True)
    try:
        with torch.no_grad():
            out = net(z["rgb"].to(DEV))["inv_depths"]
    finally:
        PF.set_im2col_first(prev)
    for i, d in enumerate(out):
        ref = z["disp%d" % (i + 1)]
        rel = ((d.cpu() - ref).abs() / ref.abs()).max().item()
        print("im2col disp%d max-rel %.3e" % (i + 1, rel))
        assert rel < 1e-3
