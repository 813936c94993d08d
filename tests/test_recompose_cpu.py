"""CPU tier: algebra of the staged re-compositions of small-channel convolutions (functional.conv2d_im2col) against the
direct convolution in float64, values and gradients.  The convolution callable is PyTorch's here; on the GPU tier it is
the tensor-core engine (tests/test_recompose_gpu.py, PN_EXPERIMENTAL=1)."""
import pytest
import torch
import torch.nn.functional as F

from packnet_sfm_b200 import functional as PF


def _torch_conv_nhwc(x, w, b):
    k = w.shape[2]
    return F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=k // 2).permute(0, 2, 3, 1)


@pytest.mark.parametrize("shape", [(2, 9, 11, 3, 6, 5), (1, 6, 7, 1, 4, 3), (1, 8, 8, 3, 5, 7), (2, 5, 6, 2, 3, 1)])
@pytest.mark.parametrize("align", [8, 4])
def test_im2col_convolution_equals_direct_convolution(shape, align):
    B, H, W, cin, cout, k = shape
    g = torch.Generator().manual_seed(B * 100 + H * 10 + k)
    x = torch.rand(B, H, W, cin + 2, generator=g, dtype=torch.float64).requires_grad_(True)    # extra channels are ignored
    w = (torch.rand(cout, cin, k, k, generator=g, dtype=torch.float64) - 0.5).requires_grad_(True)
    b = (torch.rand(cout, generator=g, dtype=torch.float64) - 0.5).requires_grad_(True)
    y = PF.conv2d_im2col(x, w, b, conv=_torch_conv_nhwc, align=align)
    gy = torch.rand(y.shape, generator=g, dtype=torch.float64) - 0.5
    y.backward(gy)
    xr, wr, br = (t.detach().clone().requires_grad_(True) for t in (x, w, b))
    yr = _torch_conv_nhwc(xr[..., :cin], wr, br)
    yr.backward(gy)
    assert y.shape == yr.shape
    assert (y - yr).abs().max() < 1e-12
    assert (x.grad - xr.grad).abs().max() < 1e-12 and (w.grad - wr.grad).abs().max() < 1e-12
    assert (b.grad - br.grad).abs().max() < 1e-12
