"""The fold and frame kernels executed on the CPU tier FROM THEIR REAL SOURCE: packnet_sfm_b200/csrc/fold_kernels.cu and
frame_kernels.cu compiled for the host with g++ -DPN_EMULATE against tests/emu/cuda_emu.h (one OS thread per CUDA thread,
std::barrier for __syncthreads, per-warp exchange for the shuffles) behind the same C-ABI entry points, then driven by
the real Python glue of packnet_sfm_b200/folded.py.  Checked against the reference composition
packing -> Conv3d -> pad -> Conv2d (layers01.py:239-247): values and every gradient.

This pins the kernels' index arithmetic, staging, synchronisation placement and host-side dispatch before they ever see
a GPU; performance and the tcgen05 convolution itself remain GPU-tier matters (tests/test_folded_gpu.py)."""
import pytest
import torch
import torch.nn.functional as F

from oracle import packnet_oracle as PO
from packnet_sfm_b200 import folded


@pytest.fixture
def kernel_path(emulated_kernels):
    return emulated_kernels


def rel(a, b):
    return float((a.detach().double() - b.detach().double()).abs().max() / b.detach().double().abs().max())


@pytest.mark.parametrize("co,C,k", [(2, 2, 3), (1, 33, 3), (1, 3, 5)])
def test_fold_kernels_emulated(kernel_path, co, C, k):
    """nine windows, both layouts, multi-block depth (n > 128), dS, accumulate, the dW3 warp/shared/global reduction"""
    n = 4 * C
    g = torch.Generator().manual_seed(co + C + k)
    w2 = (torch.rand(co, 8 * n, k, k, generator=g) - 0.5).requires_grad_(True)
    w3 = (torch.rand(8, 1, 3, 3, 3, generator=g) - 0.5).requires_grad_(True)
    outs = folded.fold_set(w2, w3)
    gs = [torch.rand(o.shape, generator=g) - 0.5 for o in outs]
    torch.autograd.backward(outs, gs)
    w2d, w3d = (t.detach().double().requires_grad_(True) for t in (w2, w3))
    refs = folded.fold_set_torch(w2d, w3d)
    torch.autograd.backward(refs, [x.double() if i in (0, 9) else x.double().permute(0, 3, 1, 2) for i, x in enumerate(gs)])
    for i, (name, a, b) in enumerate(zip(folded.FOLD_ORDER + ("S",), outs, refs)):
        if i not in (0, 9):
            b = b.permute(0, 2, 3, 1)
        assert a.shape == b.shape and rel(a, b) < 1e-5, (name, rel(a, b))
    assert rel(w2.grad, w2d.grad) < 1e-5 and rel(w3.grad, w3d.grad) < 1e-5, (rel(w2.grad, w2d.grad), rel(w3.grad, w3d.grad))


def _conv_nhwc(x, w, b):
    return F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=w.shape[-1] // 2).permute(0, 2, 3, 1).contiguous()


@pytest.mark.parametrize("case", [(2, 2, 8, 12, 3, 3), (1, 2, 10, 12, 2, 5)])
def test_folded_pack_block_with_emulated_kernels(kernel_path, case):
    """the whole kernel path of pack_conv_folded (PyTorch convolution for the O(area) part): folds, all eight frame terms,
    bias classes, both backward chains"""
    B, C, H, W, Co, k = case
    g = torch.Generator().manual_seed(B + C + H + k)
    x = (torch.rand(B, H, W, C, generator=g) - 0.5).requires_grad_(True)
    w2 = ((torch.rand(Co, 32 * C, k, k, generator=g) - 0.5) * 0.2).requires_grad_(True)
    b2 = (torch.rand(Co, generator=g) - 0.5).requires_grad_(True)
    w3 = (torch.rand(8, 1, 3, 3, 3, generator=g) - 0.5).requires_grad_(True)
    b3 = (torch.rand(8, generator=g) - 0.5).requires_grad_(True)
    z = folded.pack_conv_folded(x, w2, b2, w3, b3, _conv_nhwc)
    gz = torch.rand(z.shape, generator=g) - 0.5
    z.backward(gz)
    xd, w2d, b2d, w3d, b3d = (t.detach().double().requires_grad_(True) for t in (x, w2, b2, w3, b3))
    t = PO.conv3d_features(PO.packing(xd.permute(0, 3, 1, 2)), w3d, b3d)
    zr = F.conv2d(F.pad(t, [k // 2] * 4), w2d, b2d).permute(0, 2, 3, 1)
    zr.backward(gz.double())
    for name, a, b in (("z", z, zr), ("gx", x.grad, xd.grad), ("gw2", w2.grad, w2d.grad), ("gb2", b2.grad, b2d.grad),
                       ("gw3", w3.grad, w3d.grad), ("gb3", b3.grad, b3d.grad)):
        assert rel(a, b) < 3e-5, (name, rel(a, b))


@pytest.mark.parametrize("case", [(2, 17, 4, 35, 66, 3), (1, 3, 5, 70, 5, 5)])
def test_frame_kernels_emulated_with_several_tiles(kernel_path, case):
    """pn_pack_frame_* alone (folds from the PyTorch definition): more than one 64-wide tile in every GEMM dimension --
    pixels (B*w = 70), output columns (A*Co = 66), channels (n = 68) -- and ragged last tiles, against frame_strips."""
    B, C, h, w, Co, k = case
    n, m = 4 * C, k // 2
    g = torch.Generator().manual_seed(B + C + h + k)
    xs = torch.rand(B, h, w, n, generator=g) - 0.5
    lines = [xs[:, 0].clone(), xs[:, h - 1].clone(), xs[:, :, 0].clone(), xs[:, :, w - 1].clone()]
    w2 = (torch.rand(Co, 8 * n, k, k, generator=g) - 0.5) * 0.2
    w3 = torch.rand(8, 1, 3, 3, 3, generator=g) - 0.5
    b3 = torch.rand(8, generator=g) - 0.5
    folds = folded.fold_set_torch(w2, w3)
    gz = torch.rand(B, h, w, Co, generator=g) - 0.5
    # kernel path
    la = [t.clone().requires_grad_(True) for t in lines]
    wa = [f.permute(0, 2, 3, 1).contiguous().requires_grad_(True) for f in folds[1:9]]
    _, dB = folded.bias_classes(folds[9], b3, k, w, "cpu")
    dBa = dB.clone().requires_grad_(True)
    z0 = torch.zeros(B, h, w, Co).requires_grad_(True)
    z = folded._FrameApplyCUDA.apply(z0 * 1.0, *la, *wa, dBa, k)
    z.backward(gz)
    # definition
    lb = [t.clone().double().requires_grad_(True) for t in lines]
    fb = [f.double().clone().requires_grad_(True) for f in folds]
    b3b = b3.double().requires_grad_(True)
    _, ts, bs, ls, rs = folded.frame_strips(*lb, fb, b3b, k)
    zr = torch.zeros(B, h, w, Co, dtype=torch.float64)
    zr[:, :m] += ts
    zr[:, h - m:] += bs
    zr[:, :, :m] += ls
    zr[:, :, w - m:] += rs
    (zr * gz.double()).sum().backward()
    assert rel(z, zr) < 1e-5, rel(z, zr)
    for a, b in zip(la, lb):
        assert rel(a.grad, b.grad) < 1e-5, rel(a.grad, b.grad)
    for name, a, b in zip(folded.FOLD_ORDER[1:], wa, fb[1:9]):
        assert rel(a.grad, b.grad.permute(0, 2, 3, 1)) < 1e-5, (name, rel(a.grad, b.grad.permute(0, 2, 3, 1)))
    # bias classes: gradient of the class table = sum of gz over the pixels of each class
    from kernel_mirrors import border_class
    want = torch.zeros_like(dBa)
    for row in range(h):
        for col in range(w):
            if row < m or row >= h - m or col < m or col >= w - m:
                want[border_class(row, h, m), border_class(col, w, m)] += gz[:, row, col].sum(0)
    assert rel(dBa.grad, want) < 1e-5


def test_weight_grad_unpack_tiled_emulated(emulated_kernels):
    """pn_conv2d_unpack_weight_grad_tiled (staged shared-memory re-layout of the weight gradient) from its real source under
    the host emulation: dw[co][ci][tap] == dw_packed[co][tap][ci] bit for bit, ragged channel blocks, 1x1 .. 7x7."""
    import ctypes
    import torch
    from packnet_sfm_b200 import _lib
    lib = _lib.lib()
    torch.manual_seed(0)
    for cout, cin, k, kpad in ((3, 5, 3, 8), (2, 130, 3, 192), (4, 64, 1, 64), (1, 200, 5, 256), (2, 9, 7, 64), (1, 257, 3, 320)):
        dwp = torch.rand(cout, k * k, kpad)
        out = torch.full((cout, cin, k, k), float("nan"))
        rc = lib.pn_conv2d_unpack_weight_grad_tiled(_lib.ptr(dwp), _lib.ptr(out), cout, cin, k, kpad, None)
        assert rc == 0, lib.pn_last_error_string()
        want = dwp[:, :, :cin].permute(0, 2, 1).reshape(cout, cin, k, k)
        assert torch.equal(out, want), (cout, cin, k, kpad)


def _packed_weight_mirror(w, transposed, rows_pad):
    """the packed bf16 layout of conv_engine.cu (pack_weight_kernel), element for element, in PyTorch"""
    import torch
    cout, cin, k, _ = w.shape
    taps = k * k
    rows, kred = (cin, cout) if transposed else (cout, cin)
    cchunks = (kred + 63) // 64
    wf = w.reshape(cout, cin, taps)
    hi = torch.zeros(cchunks, taps, rows_pad, 64)
    for cc in range(cchunks):
        for r in range(rows):
            for e in range(64):
                kk = cc * 64 + (((e // 8) ^ (r & 7)) * 8 + e % 8)
                if kk >= kred:
                    continue
                hi[cc, :, r, e] = wf[kk, r].flip(0) if transposed else wf[r, kk]
    h = hi.to(torch.bfloat16)
    lo = (hi - h.float()).to(torch.bfloat16)
    return h.reshape(-1), lo.reshape(-1)


def test_weight_pack_tiled_emulated(emulated_kernels):
    """pn_conv2d_pack_weight_tiled (staged shared-memory weight packing, bf16 hi/lo, SWIZZLE_128B rows) from its real source
    under the host emulation against a PyTorch mirror of the documented layout: both orientations, ragged rows / chunks."""
    import torch
    from packnet_sfm_b200 import _lib
    lib = _lib.lib()
    torch.manual_seed(1)
    for cout, cin, k, transposed, rows_pad in ((5, 7, 3, 0, 8), (5, 7, 3, 1, 8), (16, 70, 1, 0, 16), (70, 20, 3, 1, 24), (9, 130, 5, 0, 16),
                                               (66, 9, 5, 1, 16), (3, 3, 7, 0, 8)):
        w = torch.rand(cout, cin, k, k) - 0.5
        kred = cout if transposed else cin
        n = ((kred + 63) // 64) * k * k * rows_pad * 64
        hi = torch.full((n,), float("nan")).to(torch.bfloat16)
        lo = torch.full((n,), float("nan")).to(torch.bfloat16)
        rc = lib.pn_conv2d_pack_weight_tiled(_lib.ptr(w), _lib.ptr(hi), _lib.ptr(lo), cout, cin, k, transposed, rows_pad, None)
        assert rc == 0, lib.pn_last_error_string()
        want_hi, want_lo = _packed_weight_mirror(w, bool(transposed), rows_pad)
        assert torch.equal(hi.view(torch.int16), want_hi.view(torch.int16)), (cout, cin, k, transposed)
        assert torch.equal(lo.view(torch.int16), want_lo.view(torch.int16)), (cout, cin, k, transposed)


def test_flat_adam_emulated_matches_torch_adam_and_writes_the_forward_tiles(emulated_kernels):
    """optim.FlatAdam + pn_adam_step from their real source under the host emulation: three steps against torch.optim.Adam
    on the same values (two groups with different learning rates, one stored convolution weight with ragged Cin / Cout, plain
    tensors of odd sizes), the parameter views keep shape / values / state_dict semantics, the stored weight's gradient view
    has the [Cout][tap][kpad] layout of the weight-gradient kernel, and the bf16 hi/lo forward tiles equal the mirror of the
    layout conv_engine.cu documents after every step."""
    import torch
    from packnet_sfm_b200 import optim
    torch.manual_seed(0)
    cout, cin, k = 20, 70, 3
    conv_w = torch.nn.Parameter(torch.rand(cout, cin, k, k) - 0.5)
    plain = [torch.nn.Parameter(torch.rand(*s) - 0.5) for s in ((7,), (3, 5), (1, 13, 3, 3), (2050,))]
    other = [torch.nn.Parameter(torch.rand(33) - 0.5)]
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in [conv_w] + plain + other]
    ref = torch.optim.Adam([{"params": ref_params[:5], "lr": 2e-4}, {"params": ref_params[5:], "lr": 1e-3}])
    opt = optim.FlatAdam([{"params": [conv_w] + plain, "lr": 2e-4}, {"params": other, "lr": 1e-3}], native=[conv_w])
    nat = conv_w._pn_native
    assert nat.kpad == 128 and nat.rows_pad == 64 and conv_w.shape == (cout, cin, k, k)
    assert torch.equal(conv_w.detach(), ref_params[0].detach())           # values preserved by the re-layout
    assert conv_w.stride() == (9 * 128, 1, 3 * 128, 128)

    def check_tiles():
        want_hi, want_lo = _packed_weight_mirror(conv_w.detach().contiguous(), False, nat.rows_pad)
        # the mirror packs ceil(cin/64) chunks; the stored layout has kpad/64 of them (the same here)
        assert torch.equal(nat.hi.view(torch.int16), want_hi.view(torch.int16))
        assert torch.equal(nat.lo.view(torch.int16), want_lo.view(torch.int16))

    check_tiles()
    g = torch.Generator().manual_seed(1)
    for step in range(3):
        grads = [torch.rand(p.shape, generator=g) - 0.5 for p in ref_params]
        for p, gr in zip(ref_params, grads):
            p.grad = gr.clone()
        ref.step()
        opt.zero_grad()
        # the weight-gradient kernel's output: [Cout][tap][kpad] accumulated in the flat gradient buffer
        nat.grad_flat.view(cout, k * k, nat.kpad).zero_()
        nat.grad_flat.view(cout, k * k, nat.kpad)[:, :, :cin] = grads[0].reshape(cout, cin, k * k).permute(0, 2, 1)
        assert torch.equal(conv_w.grad, grads[0])                          # ... read back through the parameter's view
        for p, gr in zip(plain + other, grads[1:]):
            p.grad = gr.clone()
        opt.step()
        for p, q in zip([conv_w] + plain + other, ref_params):
            assert torch.allclose(p.detach(), q.detach(), rtol=0, atol=2e-7), (step, float((p - q).abs().max()))
        check_tiles()
    # padding of the stored layout never moves
    pad = torch.as_strided(opt.flat_param, (cout, k * k, nat.kpad - cin), (k * k * nat.kpad, nat.kpad, 1),
                           conv_w.storage_offset() + cin)
    assert float(pad.abs().max()) == 0.0
    # writing the weights through PyTorch is noticed (version counter) and the tiles are rebuilt on demand
    with torch.no_grad():
        conv_w.mul_(0.5)
    assert conv_w._version != nat.version
    opt.repack()
    check_tiles()
    # learning-rate schedule: device-resident, takes effect without rebuilding anything
    opt.set_lr([1e-4, 5e-4])
    assert abs(float(opt.hyper[8]) - 1e-4) < 1e-10 and abs(float(opt.hyper[10]) - 5e-4) < 1e-10
