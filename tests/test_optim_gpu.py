"""GPU tier of packnet_sfm_b200.optim.FlatAdam (pn_adam_step): the flat Adam launch against torch.optim.Adam on the real
PackNet01 + PoseNet parameter set, the forward tiles it writes against the engine's own packing kernel (bit for bit), and a
stored convolution weight through functional.conv2d (forward, data gradient from the forward tiles, weight gradient
accumulated in the flat buffer) against float64.  Reference: model_wrapper.py:128-166 (Adam, lr 2e-4 per group)."""
import pytest
import torch
import torch.nn.functional as F

from conftest import rel_l2

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def test_flat_adam_three_steps_match_torch_adam_on_the_real_parameter_set():
    from packnet_sfm_b200 import functional as PF, optim
    from packnet_sfm_b200.models import SelfSupModel
    from packnet_sfm_b200.networks import native_conv_weights
    torch.manual_seed(11)
    model = SelfSupModel().to(DEV)
    ref_model = SelfSupModel().to(DEV)
    ref_model.load_state_dict(model.state_dict())
    native = native_conv_weights(model.depth_net, (192, 640))
    assert 40 < len(native) < 80
    groups = lambda m: [{"params": list(m.depth_net.parameters()), "lr": 2e-4}, {"params": list(m.pose_net.parameters()), "lr": 1e-4}]  # noqa: E731
    opt = optim.FlatAdam(groups(model), native=native)
    ref = torch.optim.Adam(groups(ref_model))
    sd = model.state_dict()
    for k, v in ref_model.state_dict().items():                       # the re-layout kept names, shapes and values
        assert v.shape == sd[k].shape and torch.equal(v, sd[k]), k
    g = torch.Generator(device=DEV).manual_seed(5)
    for step in range(3):
        opt.zero_grad()
        for p, q in zip(model.parameters(), ref_model.parameters()):
            gr = (torch.rand(p.shape, device=DEV, generator=g) - 0.5) * (10.0 ** float(torch.randint(-6, 1, (1,))))
            q.grad = gr.clone()
            if hasattr(p, "_pn_native"):
                p.grad.copy_(gr)                                        # strided view of the flat gradient buffer
            else:
                p.grad = gr.clone()
        opt.step()
        ref.step()
        torch.cuda.synchronize()
        worst = 0.0
        for (name, p), q in zip(model.named_parameters(), ref_model.parameters()):
            d = float((p.detach() - q.detach()).abs().max())
            worst = max(worst, d)
            assert d <= 3e-7, (step, name, d)
        print("step %d: max |p - p_torch| = %.2e" % (step, worst))
    # tiles written by the optimizer launch == the engine's packing kernel on the same fp32 values, bit for bit
    for p in native[:6] + native[-6:]:
        nat = p._pn_native
        w = PF._pad_channels(p.detach().contiguous(), min(nat.kpad, (p.shape[1] + 7) // 8 * 8))
        hi, lo = PF._pack_weight(w, False, PF.PRECISION_BF16X3)
        assert hi.numel() == nat.hi.numel(), (tuple(p.shape), hi.numel(), nat.hi.numel())
        assert torch.equal(hi.view(torch.int16), nat.hi.view(torch.int16)) and torch.equal(lo.view(torch.int16), nat.lo.view(torch.int16))


@pytest.mark.parametrize("case", [(2, 40, 36, 64, 48, 3), (1, 32, 24, 136, 64, 3), (4, 6, 20, 64, 128, 3), (1, 16, 16, 72, 512, 1),
                                  (2, 96, 160, 64, 64, 3), (1, 32, 24, 64, 64, 7), (1, 24, 40, 256, 32, 3)])
def test_stored_weight_through_the_engine(case):
    """A convolution weight stored by FlatAdam: forward and data gradient read the optimizer-written tiles (the data gradient
    as MN-major operands of the FORWARD tiles), the weight gradient lands in the flat buffer -- all against float64.
    Cin = 129 -> activation padded to 136 (the decoder's concatenations), Cout = 32 (rows padded to 64)."""
    from packnet_sfm_b200 import functional as PF, optim
    B, H, W, Cin_t, Cout, k = case
    cin = Cin_t - 7 if Cin_t % 64 else Cin_t          # 136 -> 129, 72 -> 65: parameters narrower than the padded activation
    torch.manual_seed(B * 100 + Cin_t + k)
    w = torch.nn.Parameter(((torch.rand(Cout, cin, k, k, device=DEV) - 0.5) * (2.0 / (cin * k * k) ** 0.5)))
    b = torch.nn.Parameter(torch.rand(Cout, device=DEV) - 0.5)
    opt = optim.FlatAdam([w, b], native=[w])
    x = torch.zeros(B, H, W, Cin_t, device=DEV)
    x[..., :cin] = torch.rand(B, H, W, cin, device=DEV) - 0.5
    x.requires_grad_(True)
    gy = torch.rand(B, H, W, Cout, device=DEV) - 0.5
    opt.zero_grad()
    y = PF.conv2d(x, w, b)
    y.backward(gy)
    torch.cuda.synchronize()
    assert w.grad.data_ptr() == opt.flat_grad.data_ptr() + 4 * w._pn_native.grad_flat.storage_offset()
    xd = x.detach()[..., :cin].double().requires_grad_(True)
    wd, bd = w.detach().double().requires_grad_(True), b.detach().double().requires_grad_(True)
    yr = F.conv2d(xd.permute(0, 3, 1, 2), wd, bd, padding=k // 2).permute(0, 2, 3, 1)
    yr.backward(gy.double())
    tol = 6e-5 + 1e-8 * max(cin * k * k, B * H * W)
    assert rel_l2(y.detach(), yr.detach()) < tol
    assert rel_l2(x.grad[..., :cin], xd.grad) < tol, rel_l2(x.grad[..., :cin], xd.grad)
    assert rel_l2(w.grad, wd.grad) < tol, rel_l2(w.grad, wd.grad)
    assert float(x.grad[..., cin:].abs().max()) < 1e-6 if Cin_t > cin else True      # zero weight columns -> zero gradient
    # one optimizer step, then the forward again: it must see the NEW weights through the tiles the step wrote
    opt.step()
    y2 = PF.conv2d(x.detach(), w, b)
    yr2 = F.conv2d(x.detach()[..., :cin].permute(0, 3, 1, 2).double(), w.detach().double(), b.detach().double(),
                   padding=k // 2).permute(0, 2, 3, 1)
    assert rel_l2(y2.detach(), yr2) < tol
    assert float((w.detach() - wd.detach().float()).abs().max()) > 1e-5     # ... and the step did move them
