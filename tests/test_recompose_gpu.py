"""GPU tier: the re-compositions and kernel variants that became defaults in round 2, each against its round-1 counterpart
or float64 -- the first-layer re-composition -- functional.conv2d_im2col on the
tensor-core engine against the direct convolution in float64, and PackNet01 with it against the reference's golden depth
maps (bar 1e-3)."""
import os

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden, rel_l2
from oracle import packnet_oracle as PO

pytestmark = pytest.mark.gpu
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


@pytest.mark.parametrize("shape", [(4, 192, 640, 64), (2, 24, 80, 256), (1, 9, 13, 16)])
def test_head_conv_flat_staging_matches_default(shape):
    """pn_set_tuning(PN_TUNE_STAGE_FLAT): the head kernels stage their tiles with all threads; the shared-memory image is
    the same, so the forward must be bit-identical and the weight gradient equal up to the order of its atomics."""
    from packnet_sfm_b200 import _lib, functional as PF
    B, H, W, C = shape
    g = torch.Generator().manual_seed(C)
    x = (torch.rand(B, H, W, C, generator=g) - 0.5).to(DEV)
    w = ((torch.rand(1, C, 3, 3, generator=g) - 0.5) * 0.2).to(DEV)
    b = (torch.rand(1, generator=g) - 0.5).to(DEV)
    gy = None
    res = []
    for flat in (0, 1):
        _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, flat)
        try:
            xs, ws, bs = (t.clone().requires_grad_(True) for t in (x, w, b))
            y = PF.head_conv(xs, ws, bs)
            if gy is None:
                gy = torch.rand(y.shape, generator=g).to(DEV) - 0.5
            y.backward(gy)
            torch.cuda.synchronize()
            res.append((y.detach().cpu(), xs.grad.cpu(), ws.grad.cpu(), bs.grad.cpu()))
        finally:
            _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, 1)    # the default
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert rel_l2(res[1][2], res[0][2]) < 1e-5 and rel_l2(res[1][3], res[0][3]) < 1e-5
    xd, wd, bd = x.cpu().double(), w.cpu().double().requires_grad_(True), b.cpu().double()
    yr = F.conv2d(xd.permute(0, 3, 1, 2), wd, bd, padding=1)[:, 0]
    yr.backward(gy.cpu().double())
    assert rel_l2(res[1][0], yr.detach()) < 1e-5 and rel_l2(res[1][2], wd.grad) < 1e-4


@pytest.mark.parametrize("pack,shape", [(True, (4, 192, 640, 64)), (False, (4, 96, 320, 32)), (False, (4, 12, 40, 256)),
                                        (True, (2, 24, 80, 128)), (False, (1, 5, 7, 24))])
def test_feature_stencils_flat_staging_matches_default(pack, shape):
    """pn_set_tuning(PN_TUNE_STAGE_FLAT) for the register-tiled feature stencils: same shared-memory image -> forward and
    input gradient bit-identical, weight / bias gradients equal up to the order of their atomics."""
    from packnet_sfm_b200 import _lib, functional as PF
    g = torch.Generator().manual_seed(shape[3])
    x = (torch.rand(*shape, generator=g) - 0.5).to(DEV)
    w3 = (torch.rand(8, 1, 3, 3, 3, generator=g) - 0.5).to(DEV)
    b3 = (torch.rand(8, generator=g) - 0.5).to(DEV)
    gy = None
    res = []
    for flat in (0, 1):
        _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, flat)
        try:
            xs, ws, bs = (t.clone().requires_grad_(True) for t in (x, w3, b3))
            y = PF.pack_features(xs, ws, bs) if pack else PF.unpack_features(xs, ws, bs)
            if gy is None:
                gy = (torch.rand(y.shape, generator=g) - 0.5).to(DEV)
            y.backward(gy)
            torch.cuda.synchronize()
            res.append((y.detach().cpu(), xs.grad.cpu(), ws.grad.cpu(), bs.grad.cpu()))
            del y, xs
        finally:
            _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, 1)    # the default
    assert torch.equal(res[0][0], res[1][0]) and torch.equal(res[0][1], res[1][1])
    assert rel_l2(res[1][2], res[0][2]) < 1e-5 and rel_l2(res[1][3], res[0][3]) < 1e-5


@pytest.mark.parametrize("shape", [(4, 192, 640, 64), (4, 48, 160, 128), (4, 12, 40, 512), (2, 6, 10, 32)])
def test_groupnorm_tree_statistics_match_default(shape):
    """pn_set_tuning(PN_TUNE_GN_TREE): shuffle reduction of the GroupNorm statistics instead of fp64 shared atomics per
    thread -- the same sums in another order (double accumulation): outputs equal to rounding, and against float64."""
    from packnet_sfm_b200 import _lib, functional as PF
    B, H, W, C = shape
    g = torch.Generator().manual_seed(C + H)
    x = (torch.rand(B, H, W, C, generator=g) * 2 - 0.7).to(DEV)
    gm, bt = (torch.rand(C, generator=g) + 0.5).to(DEV), (torch.rand(C, generator=g) - 0.5).to(DEV)
    res = []
    for tree in (0, 1):
        _lib.set_tuning(_lib.PN_TUNE_GN_TREE, tree)
        try:
            res.append(PF.groupnorm_elu(x, gm, bt, 1e-5).cpu())
            torch.cuda.synchronize()
        finally:
            _lib.set_tuning(_lib.PN_TUNE_GN_TREE, 1)    # the default
    assert rel_l2(res[1], res[0]) < 1e-6
    yr = F.elu(F.group_norm(x.cpu().double().permute(0, 3, 1, 2), 16, gm.cpu().double(), bt.cpu().double(), 1e-5)).permute(0, 2, 3, 1)
    assert rel_l2(res[1], yr) < 1e-5


@pytest.mark.parametrize("shape", [(4, 96, 320, 64), (4, 96, 320, 32), (4, 48, 160, 128), (4, 24, 80, 256), (4, 12, 40, 512), (4, 6, 20, 256),
                                   (1, 2, 3, 512), (2, 7, 9, 16), (3, 5, 11, 1024), (2, 6, 10, 48), (2, 192, 640, 64)])
@pytest.mark.parametrize("residual", [False, True])
def test_groupnorm_cluster_and_two_pass_kernels_match_float64(shape, residual):
    """GroupNorm+ELU forward and backward on the one-launch cluster kernels (statistics through distributed shared memory;
    every tensor that fits the L2) and on the two-pass kernels (PN_GN_CLUSTER=0; what the 192x640 maps run): both against
    float64 PyTorch, on the network's layer shapes at B=4 192x640 plus ragged / odd ones (1024 channels = 64-wide groups, 48
    channels = no cluster plan, one-pixel-per-CTA maps)."""
    import os
    from packnet_sfm_b200 import functional as PF
    B, H, W, C = shape
    g = torch.Generator().manual_seed(C * 7 + H)
    x0 = (torch.rand(B, H, W, C, generator=g) * 2 - 0.7).to(DEV)
    x20 = (torch.rand(B, H, W, C, generator=g) - 0.5).to(DEV)
    gm0, bt0 = (torch.rand(C, generator=g) + 0.5).to(DEV), (torch.rand(C, generator=g) - 0.5).to(DEV)
    gy = (torch.rand(B, H, W, C, generator=g) - 0.5).to(DEV)
    xd, x2d, gd, bd = (t.detach().double().requires_grad_(True) for t in (x0, x20, gm0, bt0))
    inp = xd + x2d if residual else xd
    yr = F.elu(F.group_norm(inp.permute(0, 3, 1, 2), 16, gd, bd, 1e-5)).permute(0, 2, 3, 1)
    yr.backward(gy.double())
    prev = os.environ.get("PN_GN_CLUSTER")
    try:
        for cluster in ("1", "0"):
            os.environ["PN_GN_CLUSTER"] = cluster
            x, x2, gm, bt = (t.clone().requires_grad_(True) for t in (x0, x20, gm0, bt0))
            y = PF.groupnorm_elu(x, gm, bt, 1e-5, x2=x2 if residual else None)
            y.backward(gy)
            torch.cuda.synchronize()
            assert rel_l2(y, yr) < 1e-6, (cluster, rel_l2(y, yr))
            assert rel_l2(x.grad, xd.grad) < 1e-5, (cluster, rel_l2(x.grad, xd.grad))
            assert rel_l2(gm.grad, gd.grad) < 1e-5 and rel_l2(bt.grad, bd.grad) < 1e-5, cluster
            if residual:
                assert rel_l2(x2.grad, x2d.grad) < 1e-5
            pair = getattr(y, "_pn_split", None)
            if pair is not None:      # the bf16 operand pair written next to y: hi + lo reproduces y to 16 mantissa bits
                assert rel_l2(pair[0].float() + pair[1].float(), y.detach()) < 2e-5
    finally:
        if prev is None:
            os.environ.pop("PN_GN_CLUSTER", None)
        else:
            os.environ["PN_GN_CLUSTER"] = prev


@pytest.mark.parametrize("cout,cin,k", [(64, 2048, 5), (512, 16384, 3), (64, 136, 3), (64, 64, 7), (128, 64, 1), (64, 8, 5)])
def test_weight_grad_unpack_tiled_is_bit_identical(cout, cin, k):
    """pn_conv2d_unpack_weight_grad_tiled against the default element-per-thread gather on the network's layer shapes."""
    import ctypes
    from packnet_sfm_b200 import _lib, functional as PF
    lib = _lib.lib()
    n = ctypes.c_size_t(0)
    prec = PF.get_precision()
    _lib.check(lib.pn_conv2d_wgrad_packed_elems(cout, cin, k, prec, ctypes.byref(n)), "wgrad_packed_elems")
    dwp = torch.rand(int(n.value), device=DEV)
    a = torch.empty(cout, cin, k, k, device=DEV)
    b = torch.full((cout, cin, k, k), float("nan"), device=DEV)
    st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    _lib.check(lib.pn_conv2d_unpack_weight_grad(_lib.ptr(dwp), _lib.ptr(a), cout, cin, k, prec, st), "unpack")
    _lib.check(lib.pn_conv2d_unpack_weight_grad_tiled(_lib.ptr(dwp), _lib.ptr(b), cout, cin, k, int(n.value) // (cout * k * k), st),
               "unpack tiled")
    torch.cuda.synchronize()
    assert torch.equal(a, b)


@pytest.mark.parametrize("fold", [False, True], ids=["staged_small", "staged_all"])
def test_packnet01_and_loss_with_every_staged_variant(fold):
    """bench.py --staged-small / --staged-all as a parity test: PackNet01 depth maps against the reference's golden vectors
    (bar 1e-3) and one loss forward + backward against the default kernels, with every staged switch on."""
    import ast
    from packnet_sfm_b200 import _lib, functional as PF, losses
    from packnet_sfm_b200.geometry import Pose
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    from packnet_sfm_b200.networks import PackNet01
    z = load_golden("packnet01_64x96")
    net = PackNet01(version="1A")
    net.load_state_dict(PO.packnet01_state_dict(seed=42, randomize_affine=True), strict=True)
    net = net.to(DEV).train()
    zl = load_golden("loss_fullres")
    meta = ast.literal_eval(str(zl["meta"]))

    def loss_and_grads():
        inv = [zl["inv%d" % i].to(DEV).requires_grad_(True) for i in range(meta["num_scales"])]
        mats = [zl["pose%d" % j].to(DEV).requires_grad_(True) for j in range(2)]
        K = zl["K"].to(DEV)
        out = MultiViewPhotometricLoss(**meta)(zl["rgb"].to(DEV), [zl["ctx0"].to(DEV), zl["ctx1"].to(DEV)], inv, K, K,
                                               [Pose(m) for m in mats])
        out["loss"].backward()
        torch.cuda.synchronize()
        return float(out["loss"]), [d.grad.cpu() for d in inv]

    l_ref, g_ref = loss_and_grads()
    prev = (PF.pack_fold_enabled(), PF.set_im2col_first(True), PF.set_unpack_tiled(True), losses.set_grouped_kernel(True),
            PF._state["pack_fold_min_pixels"], PF.set_pack_tiled(True))
    PF.set_pack_fold(fold, min_pixels=0)
    _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, 1)
    _lib.set_tuning(_lib.PN_TUNE_GN_TREE, 1)
    try:
        x = z["rgb"].to(DEV).requires_grad_(False)
        out = net(x)["inv_depths"]
        sum(d.sum() for d in out).backward()      # every backward kernel of the staged path runs once
        torch.cuda.synchronize()
        for i, d in enumerate(out):
            ref = z["disp%d" % (i + 1)]
            rel = ((d.detach().cpu() - ref).abs() / ref.abs()).max().item()
            print("staged disp%d max-rel %.3e" % (i + 1, rel))
            assert rel < 1e-3
        for p in net.parameters():
            assert p.grad is not None and bool(torch.isfinite(p.grad).all())
        l_new, g_new = loss_and_grads()
        assert abs(l_new - l_ref) <= 1e-5 * abs(l_ref)
        for a, b in zip(g_new, g_ref):
            err = (a.double() - b.double()).abs()
            outl = err > 1e-3 * float(b.abs().max())
            assert float(outl.double().mean()) <= max(1e-3, 21.0 / err.numel())
            assert rel_l2(a[~outl], b[~outl]) < 1e-3
    finally:
        PF.set_pack_fold(prev[0], min_pixels=prev[4])
        PF.set_im2col_first(prev[1])
        PF.set_unpack_tiled(prev[2])
        PF.set_pack_tiled(prev[5])
        losses.set_grouped_kernel(prev[3])
        _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, 1)    # the default
        _lib.set_tuning(_lib.PN_TUNE_GN_TREE, 1)    # the default


@pytest.mark.parametrize("cout,cin,k", [(64, 2048, 5), (512, 16384, 3), (64, 136, 3), (64, 64, 7), (128, 64, 1), (64, 8, 5), (136, 64, 3)])
@pytest.mark.parametrize("transposed", [False, True])
def test_weight_pack_tiled_is_bit_identical(cout, cin, k, transposed):
    """pn_conv2d_pack_weight_tiled against the default element-per-thread packing (bf16x3: hi and lo), both orientations, on
    the network's layer shapes."""
    from packnet_sfm_b200 import functional as PF
    if not PF.is_bf16(PF.get_precision()):
        pytest.skip("the tiled packing covers the bf16 precisions")
    w = (torch.rand(cout, cin, k, k, device=DEV) - 0.5) * 0.1
    prev = PF.set_pack_tiled(False)
    try:
        a_hi, a_lo = PF._pack_weight(w, transposed, PF.get_precision())
        PF.set_pack_tiled(True)
        b_hi, b_lo = PF._pack_weight(w, transposed, PF.get_precision())
        torch.cuda.synchronize()
    finally:
        PF.set_pack_tiled(prev)
    assert torch.equal(a_hi.view(torch.int16), b_hi.view(torch.int16))
    assert torch.equal(a_lo.view(torch.int16), b_lo.view(torch.int16))
