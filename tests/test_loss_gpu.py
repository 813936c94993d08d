"""GPU parity tests of the fused photometric loss (through the C-ABI via packnet_sfm_b200.losses)
against the CPU oracle and the golden vectors generated from the live reference."""
import ast

import numpy as np
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import loss_oracle as LO

import os

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True, params=["tile", "grouped", "grouped_two_launches"])
def loss_program(request):
    """Every test of this file runs on both tile programs: the grouped-scale one (csrc/loss_group_kernel.cuh,
    PN_LOSS_FLAG_GROUPED; the default since round 2) -- through the one-launch training call (pn_loss_forward_backward +
    pn_loss_backward_finish, the default) and through separate forward / backward launches -- and the first-generation one:
    same assertions."""
    from packnet_sfm_b200 import losses
    prev = losses.set_grouped_kernel(request.param != "tile")
    prev_f = losses.set_fused_training(request.param != "grouped_two_launches")
    yield "tile" if request.param == "tile" else "grouped"
    losses.set_grouped_kernel(prev)
    losses.set_fused_training(prev_f)

LOSS_TOL = 1e-3        # north_star: photometric loss within 1e-3 relative fp32
GRAD_TOL = 1e-3        # gradient fields: relative L2 over the inlier pixels
POSE_TOL = 1e-3        # pose gradients (sums over all pixels).  Measured on the B200 (r02v): 1.6e-5 .. 1.0e-4 on every fixture except
POSE_TOL_KINK = 2e-2   # loss_bigmotion (5.6e-3): there a handful of per-pixel minima resolve the other way (full-size gradient each)
POSE_TOL_ORACLE = 5e-3 # against the oracle on larger synthetic maps (configurations / bench shape: 2.2e-3 .. 2.4e-3 measured, same cause)


def _kink_pixels():
    """A winner of the per-pixel minimum that resolves the other way changes the gradient of its 3x3 neighbourhood.  fp32
    evaluates the SSIM term to ~1e-5 absolute in ANY operation order (tools/ssim_noise_study.py), so candidates closer
    than that tie; the tile program mirrors the reference's order and rarely differs (one neighbourhood allowed on tiny
    maps), the grouped program's separable sums are an independent rounding of the same noise (two neighbourhoods)."""
    from packnet_sfm_b200 import losses
    return 21.0 if losses._grouped else 12.0


def assert_field_close(got, want, tag):
    """Per-pixel gradient parity.  The loss has kinks (per-pixel min over candidates, |.|, clamp, bilinear tap
    boundaries): where two candidates tie to ~1e-7 the CUDA and CPU paths may legitimately pick different
    branches and a pixel's gradient changes by O(1).  So: at most 0.1 % such pixels (or a dozen on tiny maps),
    everything else tight."""
    got, want = got.detach().cpu().double(), want.detach().double()
    err = (got - want).abs()
    scale = float(want.abs().max()) + 1e-30
    outlier = err > 1e-3 * scale
    frac = float(outlier.double().mean())
    inl = ~outlier
    rel = float((err[inl] ** 2).sum().sqrt() / ((want[inl] ** 2).sum().sqrt() + 1e-30))
    assert frac <= max(1e-3, _kink_pixels() / err.numel()), (tag, "outlier fraction", frac)
    assert rel < GRAD_TOL, (tag, "inlier rel_l2", rel)


def _cuda_loss_from_golden(z, progress=0.0):
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    from packnet_sfm_b200.geometry import Pose
    meta = ast.literal_eval(str(z["meta"]))
    dev = torch.device("cuda:0")
    n = meta["num_scales"]
    inv = [z["inv%d" % i].to(dev).requires_grad_(True) for i in range(n)]
    mats = [z["pose%d" % j].to(dev).requires_grad_(True) for j in range(2)]
    loss_fn = MultiViewPhotometricLoss(**meta)
    K = z["K"].to(dev)
    out = loss_fn(z["rgb"].to(dev), [z["ctx0"].to(dev), z["ctx1"].to(dev)], inv, K, K, [Pose(m) for m in mats],
                  progress=progress)
    out["loss"].backward()
    torch.cuda.synchronize()
    return out, inv, mats


@pytest.mark.parametrize("case", ["loss_fullres", "loss_multires", "loss_mean_noautomask", "loss_bigmotion",
                                  "loss_progressive"])
def test_loss_and_gradients_match_reference_golden(case):
    z = load_golden(case)
    out, inv, mats = _cuda_loss_from_golden(z)
    ref = float(z["loss"])
    got = float(out["loss"].item())
    assert abs(got - ref) <= LOSS_TOL * abs(ref), (got, ref)
    assert abs(float(out["metrics"]["photometric_loss"]) - float(z["photometric_loss"])) <= LOSS_TOL * abs(ref)
    assert abs(float(out["metrics"]["smoothness_loss"]) - float(z["smoothness_loss"])) <= \
        LOSS_TOL * abs(float(z["smoothness_loss"])) + 1e-9
    assert out["loss"].shape == (1,)
    for i, d in enumerate(inv):
        g = z["ginv%d" % i]
        if d.grad is None:
            assert float(g.abs().max()) == 0.0
            continue
        assert_field_close(d.grad, g, ("ginv", i))
    for j, m in enumerate(mats):
        g = z["gpose%d" % j]
        print("POSE_REL %s gpose%d %.3e" % (case, j, rel_l2(m.grad.cpu(), g)))
        assert rel_l2(m.grad.cpu(), g) < (POSE_TOL_KINK if case == "loss_bigmotion" else POSE_TOL), ("gpose", j, rel_l2(m.grad.cpu(), g))
        assert float(m.grad[:, 3, :].abs().max()) == 0.0


@pytest.mark.parametrize("case", ["loss_fullres", "loss_bigmotion"])
def test_warp_tap_indices_bit_exact(case):
    """Integer bilinear tap origins equal the oracle's for EVERY pixel; knife-edge pixels are reported."""
    from packnet_sfm_b200.losses import warp_tap_indices
    z = load_golden(case)
    dev = torch.device("cuda:0")
    W, H = z["rgb"].shape[-1], z["rgb"].shape[-2]
    for j in range(2):
        taps, coords = warp_tap_indices(z["inv0"].to(dev), z["K"].to(dev), z["K"].to(dev), z["pose%d" % j].to(dev))
        idx, ocoords = LO.warp_tap_indices(z["inv0"], z["K"], z["K"], z["pose%d" % j])
        taps, coords = taps.cpu().numpy(), coords.cpu().numpy()
        knife = LO.knife_edge_mask(ocoords)
        print("%s ctx%d: %d px, %d knife-edge (|i-round(i)|<1e-4), coord mismatches %d, tap mismatches %d" % (
            case, j, knife.size, int(knife.sum()), int((coords != ocoords).any(-1).sum()),
            int((taps != idx).any(-1).sum())))
        assert np.array_equal(coords.view(np.int32), ocoords.view(np.int32)), "float coordinates differ bitwise"
        assert np.array_equal(taps, idx)
        # and against the coordinates the live reference produced (fixture): same integers off the knife edge
        grid = z["grid%d" % j].numpy()
        f32 = np.float32
        ix = (((grid[..., 0] + f32(1)) / f32(2)).astype(f32) * f32(W - 1)).astype(f32)
        iy = (((grid[..., 1] + f32(1)) / f32(2)).astype(f32) * f32(H - 1)).astype(f32)
        ref_idx = np.stack([np.floor(np.clip(ix, -2e9, 2e9)), np.floor(np.clip(iy, -2e9, 2e9))], -1).astype(np.int64)
        bad = (ref_idx != taps).any(-1) & ~LO.knife_edge_mask(np.stack([ix, iy], -1))
        assert not bad.any()


def _synthetic_case(B, H, W, seed=0, full_res=True):
    from packnet_sfm_b200 import synthetic
    fr = synthetic.make_frames(B, H, W, seed=100 + seed)
    inv = synthetic.make_inv_depths(B, H, W, seed=200 + seed, full_res=full_res)
    vec = synthetic.make_pose_vecs(B, seed=300 + seed)
    mats = [LO.pose_from_vec(vec[:, j]) for j in range(2)]
    return fr, inv, mats


def _run_cuda(fr, inv, mats, **kw):
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    from packnet_sfm_b200.geometry import Pose
    dev = torch.device("cuda:0")
    cfg = dict(num_scales=4, ssim_loss_weight=0.85, smooth_loss_weight=0.001, photometric_reduce_op="min",
               clip_loss=0.0, automask_loss=True)
    cfg.update(kw)
    loss_fn = MultiViewPhotometricLoss(**cfg)
    inv_d = [d.to(dev).requires_grad_(True) for d in inv]
    mats_d = [m.to(dev).requires_grad_(True) for m in mats]
    K = fr["intrinsics"].to(dev)
    out = loss_fn(fr["rgb"].to(dev), [c.to(dev) for c in fr["rgb_context"]], inv_d, K, K, [Pose(m) for m in mats_d])
    return out, inv_d, mats_d


@pytest.mark.parametrize("shape", [(1, 35, 70), (2, 48, 96), (3, 17, 33)])
def test_ragged_shapes_against_oracle(shape):
    """Tile-edge cases: sizes that are not multiples of the 32x16 tile."""
    B, H, W = shape
    fr, inv, mats = _synthetic_case(B, H, W, seed=H)
    out, inv_d, mats_d = _run_cuda(fr, inv, mats)
    out["loss"].backward()
    inv_c = [d.clone().requires_grad_(True) for d in inv]
    mats_c = [m.clone().requires_grad_(True) for m in mats]
    K = fr["intrinsics"]
    ref = LO.multiview_photometric_loss(fr["rgb"], fr["rgb_context"], inv_c, K, K, mats_c)
    ref["loss"].backward()
    assert abs(float(out["loss"].item()) - float(ref["loss"].item())) <= LOSS_TOL * abs(float(ref["loss"].item()))
    for i, (a, b) in enumerate(zip(inv_d, inv_c)):
        assert_field_close(a.grad, b.grad, ("ginv", i))
    for a, b in zip(mats_d, mats_c):
        print("POSE_REL configurations %.3e" % rel_l2(a.grad.cpu(), b.grad))
        assert rel_l2(a.grad.cpu(), b.grad) < POSE_TOL_ORACLE


def test_bench_shape_against_oracle():
    """BASELINE config[1] loss shape: B=4, 192x640, 4 full-res scales, 2 context frames."""
    B, H, W = 4, 192, 640
    fr, inv, mats = _synthetic_case(B, H, W, seed=7)
    out, inv_d, mats_d = _run_cuda(fr, inv, mats)
    out["loss"].backward()
    inv_c = [d.clone().requires_grad_(True) for d in inv]
    mats_c = [m.clone().requires_grad_(True) for m in mats]
    K = fr["intrinsics"]
    ref = LO.multiview_photometric_loss(fr["rgb"], fr["rgb_context"], inv_c, K, K, mats_c)
    ref["loss"].backward()
    got, want = float(out["loss"].item()), float(ref["loss"].item())
    print("bench-shape loss cuda %.9f oracle %.9f rel %.2e" % (got, want, abs(got - want) / abs(want)))
    assert abs(got - want) <= LOSS_TOL * abs(want)
    for i, (a, b) in enumerate(zip(inv_d, inv_c)):
        assert_field_close(a.grad, b.grad, ("ginv", i))
    for a, b in zip(mats_d, mats_c):
        print("POSE_REL bench_shape %.3e" % rel_l2(a.grad.cpu(), b.grad))
        assert rel_l2(a.grad.cpu(), b.grad) < POSE_TOL_ORACLE


def test_properties_full_size():
    """Size-independent properties at the full 384x1280 configuration."""
    B, H, W = 2, 384, 1280
    fr, inv, mats = _synthetic_case(B, H, W, seed=11)
    out1, inv1, mats1 = _run_cuda(fr, inv, mats)
    out2, _, _ = _run_cuda(fr, inv, mats)
    # deterministic loss value up to the fp64 atomic accumulation order
    assert abs(float(out1["loss"].item()) - float(out2["loss"].item())) <= 1e-6 * abs(float(out1["loss"].item()))
    # backward is linear in grad_output
    (out1["loss"] * 3.0).backward()
    g3 = [d.grad.clone() for d in inv1]
    out3, inv3, _ = _run_cuda(fr, inv, mats)
    out3["loss"].backward()
    for a, b in zip(g3, inv3):
        assert rel_l2(a, 3.0 * b.grad) < 1e-5
    # the loss is a batch mean: permuting the samples leaves it unchanged
    perm = torch.tensor([1, 0])
    frp = {"rgb": fr["rgb"][perm], "rgb_context": [c[perm] for c in fr["rgb_context"]],
           "intrinsics": fr["intrinsics"][perm]}
    outp, _, _ = _run_cuda(frp, [d[perm] for d in inv], [m[perm] for m in mats])
    assert abs(float(outp["loss"].item()) - float(out1["loss"].item())) <= 1e-5 * abs(float(out1["loss"].item()))
    # identical target and context + identity pose: the warp is the identity, warped == un-warped candidates
    eye = [torch.eye(4).repeat(B, 1, 1) for _ in range(2)]
    same = {"rgb": fr["rgb"], "rgb_context": [fr["rgb"], fr["rgb"]], "intrinsics": fr["intrinsics"]}
    outs, _, _ = _run_cuda(same, inv, eye, smooth_loss_weight=0.0)
    assert float(outs["loss"].item()) < 1e-4


def test_unsupported_options_fail_loudly():
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    fr, inv, mats = _synthetic_case(1, 32, 64)
    with pytest.raises(NotImplementedError):
        _run_cuda(fr, inv, mats, clip_loss=0.5)
    with pytest.raises(NotImplementedError):
        _run_cuda(fr, inv, mats, padding_mode="border")
    with pytest.raises(RuntimeError):
        _run_cuda(fr, inv, mats, ssim_loss_weight=0.0)
    with pytest.raises(AssertionError):
        MultiViewPhotometricLoss(automask_loss=True, photometric_reduce_op="mean")
    with pytest.raises(RuntimeError):   # CPU tensors: no fallback
        loss_fn = MultiViewPhotometricLoss(clip_loss=0.0)
        K = fr["intrinsics"]
        loss_fn(fr["rgb"], fr["rgb_context"], inv, K, K, mats)


def test_nearest_upsample_folded_into_the_kernel_matches_reference_golden(golden):
    """a8 (SfmModel.py:87-88 -> model_utils.upsample_output 'nearest'): the loss reads the four maps at their own resolution
    with index >> s.  tests/golden/loss_fullres.npz holds the LIVE reference's loss and gradients for maps that were nearest
    up-sampled to HxW, so the stored maps are its [::2^s] sub-samples, the loss must match the golden one and the gradient of
    a stored pixel is the sum of the golden gradient over its 2^s x 2^s block (the backward of the up-sampling)."""
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    from packnet_sfm_b200.geometry import Pose
    z = golden("loss_fullres")
    dev = torch.device("cuda:0")
    native = []
    for i in range(4):
        full = z["inv%d" % i]
        sub = full[..., ::1 << i, ::1 << i].contiguous()
        assert torch.equal(torch.nn.functional.interpolate(sub, full.shape[-2:], mode="nearest"), full)   # fixture sanity
        native.append(sub.to(dev).requires_grad_(True))
    mats = [z["pose%d" % j].to(dev).requires_grad_(True) for j in range(2)]
    loss_fn = MultiViewPhotometricLoss(num_scales=4, ssim_loss_weight=0.85, smooth_loss_weight=0.001, photometric_reduce_op="min",
                                       clip_loss=0.0, automask_loss=True)
    K = z["K"].to(dev)
    out = loss_fn(z["rgb"].to(dev), [z["ctx0"].to(dev), z["ctx1"].to(dev)], native, K, K, [Pose(m) for m in mats],
                  nearest_upsample=True)
    out["loss"].backward()
    want = float(z["loss"])
    assert abs(float(out["loss"].item()) - want) <= LOSS_TOL * abs(want)
    for i in range(4):
        g = z["ginv%d" % i]
        k = 1 << i
        block = torch.nn.functional.avg_pool2d(g, k) * (k * k) if i else g
        assert native[i].grad.shape == block.shape
        assert_field_close(native[i].grad, block, ("ginv", i))
    for j in range(2):
        assert rel_l2(mats[j].grad.cpu(), z["gpose%d" % j]) < POSE_TOL


def test_fused_and_explicit_upsampling_agree_at_the_bench_shape():
    """B=4, 192x640: maps at H, H/2, H/4, H/8 through the fused read against F.interpolate(nearest) + the plain kernel."""
    from packnet_sfm_b200 import synthetic
    B, H, W = 4, 192, 640
    fr, _, mats = _synthetic_case(B, H, W, seed=11)
    native = synthetic.make_inv_depths(B, H, W, seed=211, full_res=False)
    full = [torch.nn.functional.interpolate(d, (H, W), mode="nearest") if i else d for i, d in enumerate(native)]
    out_f, inv_f, mats_f = _run_cuda(fr, full, mats)
    out_f["loss"].backward()
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    from packnet_sfm_b200.geometry import Pose
    dev = torch.device("cuda:0")
    loss_fn = MultiViewPhotometricLoss(num_scales=4, ssim_loss_weight=0.85, smooth_loss_weight=0.001, photometric_reduce_op="min",
                                       clip_loss=0.0, automask_loss=True)
    inv_n = [d.to(dev).requires_grad_(True) for d in native]
    mats_n = [m.to(dev).requires_grad_(True) for m in mats]
    K = fr["intrinsics"].to(dev)
    out_n = loss_fn(fr["rgb"].to(dev), [c.to(dev) for c in fr["rgb_context"]], inv_n, K, K, [Pose(m) for m in mats_n],
                    nearest_upsample=True)
    out_n["loss"].backward()
    assert abs(float(out_n["loss"].item()) - float(out_f["loss"].item())) <= 1e-6 * abs(float(out_f["loss"].item()))
    for i in range(4):
        k = 1 << i
        want = torch.nn.functional.avg_pool2d(inv_f[i].grad, k) * (k * k) if i else inv_f[i].grad
        assert rel_l2(inv_n[i].grad, want) < 1e-5, (i, rel_l2(inv_n[i].grad, want))
    for a, b in zip(mats_n, mats_f):
        assert rel_l2(a.grad, b.grad) < 1e-4
