"""CPU-tier execution of the product's plain SIMT kernels FROM THEIR REAL SOURCE (tests/emu/: csrc/loss_kernels.cu and
csrc/layer_kernels.cu compiled for the host, one OS thread per CUDA thread), called through the product's own Python
layer (packnet_sfm_b200.losses / functional) and checked against the golden vectors of the live reference and the
oracle -- the same assertions the GPU tier makes (tests/test_loss_gpu.py, tests/test_layers_gpu.py) at sizes the
emulation finishes in seconds.  The host build rounds every float operation once (-ffp-contract=off) where nvcc may fuse
multiply-adds, so values agree to rounding; the warp-index chain is written with explicit __f*_rn intrinsics and must be
BIT-exact here as on the GPU.  The tcgen05 convolution engine has no emulation and stays a GPU-tier matter."""
import ast

import numpy as np
import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle import loss_oracle as LO
from oracle import packnet_oracle as PO

LOSS_TOL = 1e-3     # north_star bar; the emulated kernel is in fact ~1e-7 from the reference
GRAD_TOL = 1e-3


def _field_close(got, want, tag):
    """as tests/test_loss_gpu.py: a handful of kink pixels (ties of the per-pixel min, |.|, tap boundaries) may differ"""
    got, want = got.detach().double(), want.detach().double()
    err = (got - want).abs()
    scale = float(want.abs().max()) + 1e-30
    outlier = err > 1e-3 * scale
    inl = ~outlier
    rel = float((err[inl] ** 2).sum().sqrt() / ((want[inl] ** 2).sum().sqrt() + 1e-30))
    from packnet_sfm_b200 import losses
    kink = 21.0 if losses._grouped else 12.0     # see tests/test_loss_gpu.py _kink_pixels
    assert float(outlier.double().mean()) <= max(1e-3, kink / err.numel()), tag
    assert rel < GRAD_TOL, (tag, rel)


@pytest.fixture(params=[(False, True), (True, True), (True, False)], ids=["tile", "grouped", "grouped_two_launches"])
def loss_program(request):
    """both generations of the loss tile program: the first one and the grouped-scale one (PN_LOSS_FLAG_GROUPED), the latter
    through the one-launch training call (pn_loss_forward_backward + pn_loss_backward_finish) and through two launches"""
    from packnet_sfm_b200 import losses
    prev = losses.set_grouped_kernel(request.param[0])
    prev_f = losses.set_fused_training(request.param[1])
    yield request.param[0]
    losses.set_grouped_kernel(prev)
    losses.set_fused_training(prev_f)


@pytest.mark.parametrize("case", ["loss_fullres", "loss_multires", "loss_mean_noautomask", "loss_bigmotion", "loss_progressive"])
def test_emulated_loss_kernel_matches_reference_golden(emulated_kernels, loss_program, case):
    from packnet_sfm_b200.geometry import Pose
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    z = load_golden(case)
    meta = ast.literal_eval(str(z["meta"]))
    n = meta["num_scales"]
    inv = [z["inv%d" % i].clone().requires_grad_(True) for i in range(n)]
    mats = [z["pose%d" % j].clone().requires_grad_(True) for j in range(2)]
    out = MultiViewPhotometricLoss(**meta)(z["rgb"], [z["ctx0"], z["ctx1"]], inv, z["K"], z["K"], [Pose(m) for m in mats])
    out["loss"].backward()
    ref, got = float(z["loss"]), float(out["loss"].item())
    assert abs(got - ref) <= 1e-5 * abs(ref), (got, ref)
    assert abs(float(out["metrics"]["photometric_loss"]) - float(z["photometric_loss"])) <= 1e-5 * abs(ref)
    assert abs(float(out["metrics"]["smoothness_loss"]) - float(z["smoothness_loss"])) <= 1e-4 * abs(float(z["smoothness_loss"])) + 1e-9
    for i, d in enumerate(inv):
        g = z["ginv%d" % i]
        if d.grad is None:
            assert float(g.abs().max()) == 0.0
            continue
        _field_close(d.grad, g, ("ginv", i))
    for j, m in enumerate(mats):
        assert rel_l2(m.grad, z["gpose%d" % j]) < 2e-2
        assert float(m.grad[:, 3, :].abs().max()) == 0.0


@pytest.mark.parametrize("case", ["loss_fullres", "loss_bigmotion"])
def test_emulated_warp_tap_indices_bit_exact(emulated_kernels, loss_program, case):
    from packnet_sfm_b200.losses import warp_tap_indices
    z = load_golden(case)
    for j in range(2):
        taps, coords = warp_tap_indices(z["inv0"], z["K"], z["K"], z["pose%d" % j])
        idx, ocoords = LO.warp_tap_indices(z["inv0"], z["K"], z["K"], z["pose%d" % j])
        assert np.array_equal(coords.numpy().view(np.int32), ocoords.view(np.int32)), "float coordinates differ bitwise"
        assert np.array_equal(taps.numpy(), idx)


def test_emulated_loss_kernel_ragged_tiles_against_oracle(emulated_kernels, loss_program):
    """35x70 is not a multiple of the 32x16 tile: masked rows / columns, partial halos"""
    from packnet_sfm_b200 import synthetic
    from packnet_sfm_b200.geometry import Pose
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    B, H, W = 1, 35, 70
    fr = synthetic.make_frames(B, H, W, seed=135)
    inv = synthetic.make_inv_depths(B, H, W, seed=235, full_res=True)
    vec = synthetic.make_pose_vecs(B, seed=335)
    mats = [LO.pose_from_vec(vec[:, j]) for j in range(2)]
    cfg = dict(num_scales=4, ssim_loss_weight=0.85, smooth_loss_weight=0.001, photometric_reduce_op="min", clip_loss=0.0,
               automask_loss=True)
    inv_d = [d.clone().requires_grad_(True) for d in inv]
    mats_d = [m.clone().requires_grad_(True) for m in mats]
    K = fr["intrinsics"]
    out = MultiViewPhotometricLoss(**cfg)(fr["rgb"], fr["rgb_context"], inv_d, K, K, [Pose(m) for m in mats_d])
    out["loss"].backward()
    inv_c = [d.clone().requires_grad_(True) for d in inv]
    mats_c = [m.clone().requires_grad_(True) for m in mats]
    ref = LO.multiview_photometric_loss(fr["rgb"], fr["rgb_context"], inv_c, K, K, mats_c)
    ref["loss"].backward()
    assert abs(float(out["loss"].item()) - float(ref["loss"].item())) <= 1e-5 * abs(float(ref["loss"].item()))
    for i, (a, b) in enumerate(zip(inv_d, inv_c)):
        _field_close(a.grad, b.grad, ("ginv", i))
    for a, b in zip(mats_d, mats_c):
        assert rel_l2(a.grad, b.grad) < 2e-2


@pytest.mark.parametrize("flat", [0, 1], ids=["stage_tile", "stage_tile_flat"])
def test_emulated_feature_stencils(emulated_kernels, flat):
    """Conv3d(1->8) feature stencils fused with space-to-depth / depth-to-space: the register-tiled 8-depth kernels, the
    generic kernels (depth not a multiple of 8), masked rows / columns; flat=1: the staged all-threads tile staging
    (pn_set_tuning PN_TUNE_STAGE_FLAT) of the register-tiled kernels"""
    from packnet_sfm_b200 import _lib, functional as PF
    _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, flat)
    try:
        _feature_stencil_cases(PF)
    finally:
        _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, 0)


def _feature_stencil_cases(PF):
    torch.manual_seed(3)
    for pack, shape in ((True, (1, 8, 12, 8)), (False, (1, 5, 7, 24)), (True, (1, 6, 10, 5)), (False, (2, 3, 4, 16)),
                        (False, (1, 4, 37, 32)), (True, (1, 6, 70, 16))):     # several w tiles, D = 32 / 64
        x = (torch.rand(*shape) - 0.5).requires_grad_(True)
        w3 = (torch.rand(8, 1, 3, 3, 3) - 0.5).requires_grad_(True)
        b3 = (torch.rand(8) - 0.5).requires_grad_(True)
        y = PF.pack_features(x, w3, b3) if pack else PF.unpack_features(x, w3, b3)
        gy = torch.rand_like(y) - 0.5
        if shape[-1] % 4 == 0 and not pack:
            # as in the decoder: the output is concatenated with a skip tensor, so its gradient arrives as a channel WINDOW
            # of the concatenation's gradient (a narrow() view), which the kernels read in place through their channel stride
            skip = torch.rand(*y.shape[:3], 8 + y.shape[3])
            gcat = torch.cat([gy, torch.rand_like(skip)], -1)
            (torch.cat([y, skip], -1) * gcat).sum().backward()
        else:
            y.backward(gy)
        xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w3, b3))
        xn = xd.permute(0, 3, 1, 2)
        if pack:
            yr = PO.conv3d_features(PO.packing(xn), wd, bd).permute(0, 2, 3, 1)
        else:
            yr = F.pixel_shuffle(PO.conv3d_features(xn, wd, bd), 2).permute(0, 2, 3, 1)
        yr.backward(gy.double())
        assert rel_l2(y, yr) < 1e-6, (pack, shape)
        assert rel_l2(x.grad, xd.grad) < 1e-5 and rel_l2(w3.grad, wd.grad) < 1e-4 and rel_l2(b3.grad, bd.grad) < 1e-4, (pack, shape)


def test_emulated_groupnorm_elu_and_head_conv(emulated_kernels):
    from packnet_sfm_b200 import _lib, functional as PF
    torch.manual_seed(4)
    # (C, tree): tree = the staged shuffle reduction of the statistics kernel (pn_set_tuning PN_TUNE_GN_TREE): one float4
    # column per group (64), two (128), four columns over two warps per pixel (256); 16 and 32 fall back to the atomics
    # (80, 120): 65 pixels per CTA -> the four-pixel trips of the TREE kernel also run at C = 64 (16 pixel lanes)
    for C, tree, hw in ((16, 0, (6, 10)), (64, 0, (6, 10)), (16, 1, (6, 10)), (32, 1, (6, 10)), (64, 1, (6, 10)), (128, 1, (6, 10)),
                        (256, 1, (6, 10)), (64, 1, (80, 120))):
        _lib.set_tuning(_lib.PN_TUNE_GN_TREE, tree)
        x = (torch.rand(2, hw[0], hw[1], C) * 2 - 0.7).requires_grad_(True)
        x2 = (torch.rand(2, hw[0], hw[1], C) - 0.5).requires_grad_(True)
        g = (torch.rand(C) + 0.5).requires_grad_(True)
        bt = (torch.rand(C) - 0.5).requires_grad_(True)
        for second in (None, x2):
            for t in (x, x2, g, bt):
                t.grad = None
            y = PF.groupnorm_elu(x, g, bt, 1e-5, x2=second)
            gy = torch.rand_like(y) - 0.5
            # the bf16 operand pair the apply pass wrote for the consuming convolution: hi = rn(y), lo = rn(y - hi), bit for bit
            hi, lo, ver = y._pn_split
            assert ver == y._version and torch.equal(hi.view(torch.int16), y.detach().to(torch.bfloat16).view(torch.int16))
            assert torch.equal(lo.view(torch.int16), (y.detach() - hi.float()).to(torch.bfloat16).view(torch.int16))
            assert PF._operands(y, PF.PRECISION_BF16X3)[0] is hi
            seen = {}
            def hook(gr):
                seen["dx"] = getattr(gr, "_pn_split", None)
            h = x.register_hook(hook)
            y.backward(gy)
            h.remove()
            if seen.get("dx") is not None:       # ... and of dx in the backward (the attribute travels with the tensor object)
                dh, dl, _ = seen["dx"]
                assert torch.equal(dh.view(torch.int16), x.grad.to(torch.bfloat16).view(torch.int16))
                assert torch.equal(dl.view(torch.int16), (x.grad - dh.float()).to(torch.bfloat16).view(torch.int16))
            xd, x2d, gd, bd = (t.detach().double().requires_grad_(True) for t in (x, x2, g, bt))
            inp = xd if second is None else xd + x2d
            yr = F.elu(F.group_norm(inp.permute(0, 3, 1, 2), 16, gd, bd, 1e-5)).permute(0, 2, 3, 1)
            yr.backward(gy.double())
            assert rel_l2(y, yr) < 1e-6
            assert rel_l2(x.grad, xd.grad) < 1e-5 and rel_l2(g.grad, gd.grad) < 1e-5 and rel_l2(bt.grad, bd.grad) < 1e-5
            if second is not None:
                assert rel_l2(x2.grad, x2d.grad) < 1e-5
    _lib.set_tuning(_lib.PN_TUNE_GN_TREE, 0)
    results = {}
    for flat in (0, 1):      # 1: the staged all-threads tile staging of the head kernels (pn_set_tuning PN_TUNE_STAGE_FLAT)
        _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, flat)
        try:
            torch.manual_seed(5)
            for B, H, W, C in ((1, 9, 13, 16), (2, 6, 20, 8), (1, 17, 40, 64)):
                x = (torch.rand(B, H, W, C) - 0.5).requires_grad_(True)
                w = ((torch.rand(1, C, 3, 3) - 0.5) * 0.2).requires_grad_(True)
                b = (torch.rand(1) - 0.5).requires_grad_(True)
                y = PF.head_conv(x, w, b)
                gy = torch.rand_like(y) - 0.5
                y.backward(gy)
                xd, wd, bd = (t.detach().double().requires_grad_(True) for t in (x, w, b))
                yr = F.conv2d(xd.permute(0, 3, 1, 2), wd, bd, padding=1)[:, 0]
                yr.backward(gy.double())
                assert rel_l2(y, yr) < 1e-6 and rel_l2(x.grad, xd.grad) < 1e-6
                assert rel_l2(w.grad, wd.grad) < 1e-5 and rel_l2(b.grad, bd.grad) < 1e-5
                results.setdefault((B, H, W, C), []).append((y.detach().clone(), w.grad.clone()))
        finally:
            _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, 0)
    for (y0, w0), (y1, w1) in results.values():
        assert torch.equal(y0, y1)      # the same shared-memory image -> the forward is bit-identical
        assert rel_l2(w1, w0) < 1e-6    # weight gradient: atomics, order may differ


@pytest.mark.parametrize("B,H,W,N,full_res,kw", [
    (1, 16, 32, 1, True, {}),                                  # one context frame
    (1, 16, 32, 3, True, {}),                                  # three (the reference's configs use two)
    (2, 8, 8, 2, True, {}),                                    # a map smaller than one 32x16 tile
    (1, 20, 40, 2, True, {"smooth_loss_weight": 0.0}),
    (1, 20, 40, 2, True, {"num_scales": 2}),
    (1, 32, 64, 2, False, {"photometric_reduce_op": "mean", "automask_loss": False}),   # multi-resolution scales + mean
], ids=["N1", "N3", "tiny", "nosmooth", "2scales", "multires_mean"])
def test_emulated_loss_kernel_configurations_against_oracle(emulated_kernels, B, H, W, N, full_res, kw):
    """template instantiations and switches the golden vectors do not reach (context count, scale count, reductions)"""
    from packnet_sfm_b200 import synthetic
    from packnet_sfm_b200.geometry import Pose
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    fr = synthetic.make_frames(B, H, W, seed=100 + N)
    ctx = list(fr["rgb_context"])
    while len(ctx) < N:
        ctx.append(torch.roll(fr["rgb"], 2 + len(ctx), 3) * 0.9 + 0.05)
    ctx = ctx[:N]
    inv = synthetic.make_inv_depths(B, H, W, seed=200 + N, full_res=full_res)
    g = torch.Generator().manual_seed(N)
    scale = torch.tensor([0.2, 0.2, 0.2, 0.02, 0.02, 0.02])
    mats = [LO.pose_from_vec((torch.rand(B, 6, generator=g) - 0.5) * scale) for _ in range(N)]
    cfg = dict(num_scales=4, ssim_loss_weight=0.85, smooth_loss_weight=0.001, photometric_reduce_op="min", clip_loss=0.0,
               automask_loss=True)
    cfg.update(kw)
    K = fr["intrinsics"]
    inv_d = [d.clone().requires_grad_(True) for d in inv]
    mats_d = [m.clone().requires_grad_(True) for m in mats]
    out = MultiViewPhotometricLoss(**cfg)(fr["rgb"], ctx, inv_d, K, K, [Pose(m) for m in mats_d])
    out["loss"].backward()
    inv_c = [d.clone().requires_grad_(True) for d in inv]
    mats_c = [m.clone().requires_grad_(True) for m in mats]
    okw = {k: v for k, v in cfg.items() if k != "clip_loss"}
    ref = LO.multiview_photometric_loss(fr["rgb"], ctx, inv_c, K, K, mats_c, **okw)
    ref["loss"].backward()
    assert abs(float(out["loss"].item()) - float(ref["loss"].item())) <= 1e-5 * abs(float(ref["loss"].item()))
    for i in range(cfg["num_scales"]):
        _field_close(inv_d[i].grad, inv_c[i].grad, ("ginv", i))
    for a, b in zip(mats_d, mats_c):
        assert rel_l2(a.grad, b.grad) < 2e-2


def test_emulated_packnet01_forward_matches_reference_golden():
    """tests/emu/packnet01_emulated.py: the product's PackNet01 (networks.py) with every SIMT kernel from its real source
    under the emulation and PyTorch's conv2d in place of the tcgen05 engine, against the golden depth maps of the live
    reference.  Runs as a script in its own process (it swaps functional.conv2d and lifts the CUDA-only guard)."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import __graft_entry__ as ge
    ge._build_kernel_emulation()
    r = subprocess.run([sys.executable, os.path.join(root, "tests", "emu", "packnet01_emulated.py")], stdout=subprocess.PIPE,
                       stderr=subprocess.STDOUT, text=True, timeout=500)
    assert r.returncode == 0, r.stdout[-2000:]
    assert "disp4 max-rel" in r.stdout


def test_emulated_loss_reads_the_maps_nearest_upsampled(emulated_kernels, loss_program):
    """a8 folded into the kernel (pn_loss_desc.inv_shift): the four maps at their own resolution, read with index >> s, against the
    LIVE reference's golden loss / gradients for the nearest up-sampled maps (tests/golden/loss_fullres.npz): same loss, and the
    gradient of a stored pixel = the golden gradient summed over its 2^s x 2^s block."""
    import torch.nn.functional as F
    from packnet_sfm_b200.geometry import Pose
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    z = load_golden("loss_fullres")
    meta = ast.literal_eval(str(z["meta"]))
    native = []
    for i in range(4):
        full = z["inv%d" % i]
        sub = full[..., ::1 << i, ::1 << i].contiguous()
        assert torch.equal(F.interpolate(sub, full.shape[-2:], mode="nearest"), full)
        native.append(sub.requires_grad_(True))
    mats = [z["pose%d" % j].clone().requires_grad_(True) for j in range(2)]
    out = MultiViewPhotometricLoss(**meta)(z["rgb"], [z["ctx0"], z["ctx1"]], native, z["K"], z["K"], [Pose(m) for m in mats],
                                           nearest_upsample=True)
    out["loss"].backward()
    ref, got = float(z["loss"]), float(out["loss"].item())
    assert abs(got - ref) <= 1e-5 * abs(ref), (got, ref)
    for i in range(4):
        k = 1 << i
        want = F.avg_pool2d(z["ginv%d" % i], k) * (k * k) if i else z["ginv%d" % i]
        assert native[i].grad.shape == want.shape
        _field_close(native[i].grad, want, ("ginv", i))
    for j, m in enumerate(mats):
        assert rel_l2(m.grad, z["gpose%d" % j]) < 2e-2
