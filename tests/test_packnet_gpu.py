"""GPU parity of the whole PackNet01 depth network against the golden vectors from the live reference and
against the CPU oracle (depth maps within 1e-3 relative, north_star)."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import packnet_oracle as PO

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.fixture(params=["default", "fold_all", "unfolded"])
def pack_path(request):
    """default = pack1..3 folded into one (k+2)x(k+2) convolution where the map has >= 1920 pixels (policy of round 2);
    unfolded = feature stencil + convolution over the 8x-inflated channel count everywhere (the round-1 path, which still
    serves pack4 / pack5)."""
    from packnet_sfm_b200 import functional as PF
    if request.param == "unfolded":
        PF.set_pack_fold(False)
    elif request.param == "fold_all":      # every pack layer whose map is larger than the frame (also the small test maps)
        PF.set_pack_fold(True, min_pixels=0)
    yield request.param
    PF.set_pack_fold(True, min_pixels=1920)


def _net(sd):
    from packnet_sfm_b200.networks import PackNet01
    net = PackNet01(version="1A")
    net.load_state_dict(sd, strict=True)
    return net.to(DEV).train()


def test_packnet01_depth_maps_match_reference_golden(pack_path):
    z = load_golden("packnet01_64x96")
    net = _net(PO.packnet01_state_dict(seed=42, randomize_affine=True))
    with torch.no_grad():
        out = net(z["rgb"].to(DEV))["inv_depths"]
    for i, d in enumerate(out):
        ref = z["disp%d" % (i + 1)]
        rel = ((d.cpu() - ref).abs() / ref.abs()).max().item()
        print("disp%d max-rel %.3e rel-l2 %.3e" % (i + 1, rel, rel_l2(d.cpu(), ref)))
        assert d.shape == ref.shape
        assert rel < 1e-3


def test_packnet01_other_precisions():
    """tf32x3 (22-bit split) is tighter than the default bf16x3; tf32x1 (cuDNN's default on the reference's GPUs)
    misses the 1e-3 bar but stays sane."""
    from packnet_sfm_b200 import functional as PF
    z = load_golden("packnet01_64x96")
    net = _net(PO.packnet01_state_dict(seed=42, randomize_affine=True))
    try:
        for prec, name, bound in ((PF.PRECISION_TF32X3, "tf32x3", 1e-3), (PF.PRECISION_TF32X1, "tf32x1", 2e-1)):
            PF.set_precision(prec)
            with torch.no_grad():
                out = net(z["rgb"].to(DEV))["inv_depths"]
            for i, d in enumerate(out):
                ref = z["disp%d" % (i + 1)]
                rel = ((d.cpu() - ref).abs() / ref.abs()).max().item()
                print("%s disp%d max-rel %.3e rel-l2 %.3e" % (name, i + 1, rel, rel_l2(d.cpu(), ref)))
                assert rel < bound
    finally:
        PF.set_precision(PF.PRECISION_BF16X3)


def test_packnet01_gradients_match_oracle_autograd(pack_path):
    """Full backward through every custom kernel against CPU autograd of the oracle restatement."""
    from packnet_sfm_b200 import synthetic
    sd = PO.packnet01_state_dict(seed=7, randomize_affine=True)
    net = _net(sd)
    x = synthetic.make_frames(1, 64, 96, seed=9)["rgb"]
    out = net(x.to(DEV))["inv_depths"]
    g = torch.Generator().manual_seed(1)
    gouts = [torch.rand(d.shape, generator=g) - 0.5 for d in out]
    torch.autograd.backward(out, [t.to(DEV) for t in gouts])
    torch.cuda.synchronize()
    sdc = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    ref = PO.packnet01_forward(x, sdc)
    torch.autograd.backward(ref, gouts)
    for a, b in zip(out, ref):
        assert rel_l2(a.detach().cpu(), b.detach()) < 1e-3
    worst = ("", 0.0)
    for k, p in net.named_parameters():
        r = sdc[k].grad
        err = float((p.grad.cpu().double() - r.double()).norm())
        bound = 2e-3 * float(r.double().norm()) + 1e-5 * r.numel() ** 0.5
        if err / max(bound, 1e-30) > worst[1]:
            worst = (k, err / max(bound, 1e-30))
        assert err <= bound, (k, err, bound)
    print("worst parameter-gradient error/bound: %s %.3f" % worst)


def test_eval_mode_contract_and_shape_errors():
    net = _net(PO.packnet01_state_dict(seed=1)).eval()
    with torch.no_grad():
        out = net(torch.rand(1, 3, 32, 64, device=DEV))
    assert torch.is_tensor(out["inv_depths"]) and out["inv_depths"].shape == (1, 1, 32, 64)
    with pytest.raises(ValueError):
        net(torch.rand(1, 3, 30, 64, device=DEV))


def strided_index(numel, n):
    """the sample positions of oracle/gen_golden.py::strided_index"""
    n = min(n, numel)
    return (torch.arange(n, dtype=torch.int64) * (numel - 1)) // max(n - 1, 1)


@pytest.mark.parametrize("fixture,B,H,W", [("packnet01_192x640_b4", 4, 192, 640), ("packnet01_384x1280_b2", 2, 384, 1280)])
def test_packnet01_baseline_config_matches_reference_golden(pack_path, fixture, B, H, W):
    """BASELINE configs[1] (B=4, 192x640) and the per-GPU shape of configs[2] (B=2, 384x1280; full-resolution map stored at
    every second pixel): the engine paths the 64x96 fixture does not reach -- persistent tile loop with
    hundreds of work items, batch folding on the small maps, the split-K thresholds -- against the LIVE reference's depth
    maps (tests/golden/packnet01_192x640_b4.npz: fp32 maps in full, gradients of sum_i <disp_i, gy_i> as per-parameter
    norms + 64 strided samples).  Bar: depth 1e-3 max-relative (north_star); gradients 2e-3 of the tensor's norm."""
    from packnet_sfm_b200 import synthetic
    z = load_golden(fixture)
    x = synthetic.make_frames(B, H, W, seed=int(z["seed_rgb"]))["rgb"]
    assert abs(float(x.double().sum()) - float(z["rgb_sum"])) < 1e-6 * float(z["rgb_sum"])       # same synthetic input
    assert torch.equal(x.reshape(-1)[::100003], z["rgb_probe"])
    net = _net(PO.packnet01_state_dict(seed=42, randomize_affine=True))
    out = net(x.to(DEV))["inv_depths"]
    g = torch.Generator().manual_seed(int(z["seed_gy"]))
    gys = [torch.rand(d.shape, generator=g) - 0.5 for d in out]
    worst = 0.0
    stride1 = int(z["disp1_stride"]) if "disp1_stride" in z else 1
    for i, d in enumerate(out):
        ref = z["disp%d" % (i + 1)]
        got = d.detach().cpu()
        if i == 0 and stride1 > 1:
            got = got[..., ::stride1, ::stride1]
        assert got.shape == ref.shape
        rel = ((got - ref).abs() / ref.abs()).max().item()
        worst = max(worst, rel)
        print("disp%d %s max-rel %.3e rel-l2 %.3e" % (i + 1, tuple(d.shape), rel, rel_l2(got, ref)))
    assert worst < 1e-3
    torch.autograd.backward(out, [t.to(DEV) for t in gys])
    torch.cuda.synchronize()
    worst_g = ("", 0.0)
    for k, p in net.named_parameters():
        norm_ref = float(z["gnorm/" + k])
        samp_ref = z["gsamp/" + k].double()
        flat = p.grad.reshape(-1)
        got = flat[strided_index(flat.numel(), 64).to(DEV)].cpu().double()
        norm_got = float(flat.double().norm())
        # a sample of n elements carries about sqrt(n / numel) of the tensor's norm: scale the bound the same way
        bound = 5e-3 * float(samp_ref.norm()) + 2e-3 * norm_ref * (got.numel() / flat.numel()) ** 0.5 + 1e-7
        err = float((got - samp_ref).norm())
        score = max(err / bound, abs(norm_got - norm_ref) / (2e-3 * norm_ref + 1e-6))
        if score > worst_g[1]:
            worst_g = (k, score)
        assert abs(norm_got - norm_ref) <= 2e-3 * norm_ref + 1e-6, (k, norm_got, norm_ref)
        assert err <= bound, (k, err, bound)
    print("worst parameter-gradient error/bound at B=%d %dx%d: %s %.3f" % ((B, H, W) + worst_g))
