"""Folded pack block: the two linear stages of PackLayerConv3d (layers01.py:239-247) as ONE convolution.

    reference:  xs = packing(x)                      [B, n=4C, h, w]        (layers01.py:126-148, 243)
                T  = Conv3d(1->8, 3x3x3, pad 1)(xs)  [B, 8n,  h, w]         (layers01.py:236-237, 244-246)
                z  = Conv2d(8n -> Co, k x k)(zero-pad k//2 (T))             (layers01.py:28-30, 36, 247)

Nothing but the zero padding sits between the Conv3d and the Conv2d, so z is linear in xs and, away from the border,
equals a single Conv2d(n -> Co, (k+2) x (k+2)) of xs whose weight is the FULL convolution of the two kernels:

    W_eff[co, c'', ey, ex] = sum_{f, dc, dy, dx}  W2[co, f*n + (c''-dc+1), ey-dy, ex-dx] * W3[f, dc, dy, dx]

(= conv_transpose3d of W2 viewed as [Co, 8, n, k, k] with W3; the depth index c'' is clipped to [0, n) exactly as the
Conv3d's zero padding in depth does).  The reduction length drops from 8n*k*k to n*(k+2)^2 -- pack1: 51 200 -> 12 544
(4.08x fewer MACs), pack2..5: 288C -> 100C (2.88x) -- and the [B, 8n, h, w] intermediate (1.0 GB for pack1 at B=4,
192x640), its bf16 split and the pack feature-stencil kernels disappear from the step.

What the single convolution gets wrong is the FRAME of width m = k//2: the reference zero-pads T, while the folded
kernel implicitly continues the Conv3d one pixel outside the map (the ring q at distance 1, where T_full(q) != 0 because
its 3x3 footprint still touches the border row / column of xs) and keeps the Conv3d bias there.  Both are repaired
exactly with thin strips:

    z = conv(xs, W_eff) + b2 + beta          (beta = interior value of the Conv3d-bias term)
        - top - bottom - left - right        (ring row / column folded into 1 x (k+2) resp. (k+2) x 1 kernels that
                                              read only the border row / column of xs)
        + 4 corner blocks                    (the ring corners are in a row AND a column term: inclusion-exclusion)
        + dB                                 (bias term on the frame: fewer taps of W2 see an in-map T)

The O(area) convolution runs on the tcgen05 engine (functional.conv2d); the nine weight folds (interior, four sides,
four corners) and their gradients on the fold kernels (csrc/fold_kernels.cu, pn_pack_fold_*); the O(perimeter) frame
terms -- eight small fp32 GEMMs plus the bias classes -- in one launch forward and three backward
(csrc/frame_kernels.cu, pn_pack_frame_*; exact fp32 FMA, no library call, so neither cuDNN's nor cuBLAS's TF32 switches
can touch the border pixels).  The s2d tensor uses the channel order (i, j, c) -- 2C contiguous floats of the NHWC source
per half-row -- instead of the reference's (c, i, j); the folded weights are permuted accordingly, so no tensor in the
reference's channel order is ever materialised.  frame_strips() / fold_set_torch() state the same terms with PyTorch ops:
they are the definition the kernels are tested against and what the CPU algebra tests run.

Status (round 1): algebra verified on the CPU against the reference composition in float64 (tests/test_folded_cpu.py:
values and all gradients to 1e-11); the kernels' index arithmetic verified against line-by-line Python mirrors
(tests/test_fold_mirror_cpu.py, tests/test_frame_mirror_cpu.py); NOT yet run on the B200 (the round's GPU budget was
spent) -- PackLayerConv3d uses it only when functional.set_pack_fold(True) / PN_PACK_FOLD=1 is given, and its GPU tests
(tests/test_folded_gpu.py) only run with PN_EXPERIMENTAL=1."""
import ctypes

import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------
# space-to-depth in the (i, j, c) channel order, with the four border rows / columns as separate small outputs
# ---------------------------------------------------------------------------------------------------------------
class _SpaceToDepthBorders(torch.autograd.Function):
    """x [B,2h,2w,C] NHWC -> xs [B,h,w,4C] with xs[b,y,x,(i*2+j)*C+c] = x[b,2y+i,2x+j,c] (packing(), layers01.py:126-148,
    up to the channel order), plus copies of its first / last row and column (the only data the frame terms read).
    The backward scatters the border gradients into the depth-to-space result instead of materialising four
    full-size zero tensors (what slicing xs under autograd would do)."""

    @staticmethod
    def forward(ctx, x):
        B, H, W, C = x.shape
        h, w = H // 2, W // 2
        xs = x.view(B, h, 2, w, 2, C).permute(0, 1, 3, 2, 4, 5).reshape(B, h, w, 4 * C)
        ctx.dims = (B, h, w, C)
        return xs, xs[:, 0].clone(), xs[:, h - 1].clone(), xs[:, :, 0].clone(), xs[:, :, w - 1].clone()

    @staticmethod
    def backward(ctx, g_xs, g_top, g_bot, g_left, g_right):
        B, h, w, C = ctx.dims
        if g_xs is None:
            gx6 = torch.zeros(B, h, 2, w, 2, C, dtype=g_top.dtype, device=g_top.device)
        else:
            gx6 = g_xs.reshape(B, h, w, 2, 2, C).permute(0, 1, 3, 2, 4, 5).contiguous()     # [B,h,i,w,j,C], fresh tensor
        if g_top is not None:
            gx6[:, 0] += g_top.reshape(B, w, 2, 2, C).permute(0, 2, 1, 3, 4)
        if g_bot is not None:
            gx6[:, h - 1] += g_bot.reshape(B, w, 2, 2, C).permute(0, 2, 1, 3, 4)
        if g_left is not None:
            gx6[:, :, :, 0] += g_left.reshape(B, h, 2, 2, C)
        if g_right is not None:
            gx6[:, :, :, w - 1] += g_right.reshape(B, h, 2, 2, C)
        return gx6.view(B, 2 * h, 2 * w, C)


def space_to_depth_borders(x):
    return _SpaceToDepthBorders.apply(x)


# ---------------------------------------------------------------------------------------------------------------
# the nine weight folds
# ---------------------------------------------------------------------------------------------------------------
def fold_windows(k):
    """(tap window of W2, face of W3) of every fold: name -> ((ky0,ky1),(kx0,kx1),(dy0,dy1),(dx0,dx1)).
    A ring row above the map is reached by the first m rows of taps and touches the map through the LAST row of the
    Conv3d kernel (dy = 2); below, by the last m rows through dy = 0; columns alike."""
    m = k // 2
    lo, hi, al, f3 = (0, m), (m + 1, k), (0, k), (0, 3)
    return {"main": (al, al, f3, f3),
            "top": (lo, al, (2, 3), f3), "bottom": (hi, al, (0, 1), f3),
            "left": (al, lo, f3, (2, 3)), "right": (al, hi, f3, (0, 1)),
            "tl": (lo, lo, (2, 3), (2, 3)), "tr": (lo, hi, (2, 3), (0, 1)),
            "bl": (hi, lo, (0, 1), (2, 3)), "br": (hi, hi, (0, 1), (0, 1))}


FOLD_ORDER = ("main", "top", "bottom", "left", "right", "tl", "tr", "bl", "br")


def _ct3(w2s, w3s):
    """full convolution of a W2 slab [Co, 8, n, a, b] with a W3 slab [8, 1, 3, c, d] -> [Co, n, a+c-1, b+d-1]
    (depth clipped to [0, n): the Conv3d pads the depth with zeros, those taps never meet data)."""
    return F.conv_transpose3d(w2s.contiguous(), w3s.contiguous(), padding=(1, 0, 0)).squeeze(1)


def _perm_n(t):
    """folded weight [Co, n=(c,i,j), ...] in the reference's s2d channel order -> ours, (i,j,c)."""
    co, n = t.shape[:2]
    return t.reshape(co, n // 4, 4, *t.shape[2:]).transpose(1, 2).reshape(co, n, *t.shape[2:])


def fold_set_torch(w2, w3):
    """The nine folds and S = sum over depth of W2 with PyTorch ops (the definition; CPU tests, GPU cross-check)."""
    co, c8, k, _ = w2.shape
    w2r = w2.reshape(co, 8, c8 // 8, k, k)
    out = []
    for name in FOLD_ORDER:
        ky, kx, dy, dx = fold_windows(k)[name]
        out.append(_perm_n(_ct3(w2r[:, :, :, ky[0]:ky[1], kx[0]:kx[1]], w3[:, :, :, dy[0]:dy[1], dx[0]:dx[1]])))
    return tuple(out) + (w2r.sum(2),)


class _FoldSetCUDA(torch.autograd.Function):
    """fold_set_torch on the fold kernels, one C call each way (pn_pack_fold_set_forward / _backward: nine launches;
    the backward accumulates all nine gradients into ONE dW2 -- the interior fold writes every tap, the border folds add
    into their windows -- and one dW3).  The interior weight is OIHW (what pn_conv2d_pack_weight takes), the eight frame
    weights channels-last [Co, EA, EB, n] (the reduction index of the frame GEMMs contiguous)."""

    @staticmethod
    def forward(ctx, w2, w3):
        from . import _lib
        _lib.require_f32(w2, w3)
        co, c8, k, _ = w2.shape
        n = c8 // 8
        m = k // 2
        w2c, w3c = w2.detach().contiguous(), w3.detach().contiguous()
        dev = w2.device
        shapes = [(co, n, k + 2, k + 2), (co, m, k + 2, n), (co, m, k + 2, n), (co, k + 2, m, n), (co, k + 2, m, n)] + \
                 [(co, m, m, n)] * 4
        outs = [torch.empty(sh, dtype=torch.float32, device=dev) for sh in shapes]
        _lib.check(_lib.lib().pn_pack_fold_set_forward(co, n, k, _lib.ptr(w2c), _lib.ptr(w3c), _lib.ptr_array(outs),
                                                       _lib.current_stream()), "pn_pack_fold_set_forward")
        S = w2c.view(co, 8, n, k, k).sum(2)
        ctx.save_for_backward(w2c, w3c)
        return tuple(outs) + (S,)

    @staticmethod
    def backward(ctx, *grads):
        from . import _lib
        w2c, w3c = ctx.saved_tensors
        co, c8, k, _ = w2c.shape
        n = c8 // 8
        dev = w2c.device
        from . import functional as PF
        PF.wait_for_wgrad_stream(*grads)      # the interior fold's gradient may come from the weight-gradient side stream
        dw2 = torch.empty_like(w2c)
        dw3 = torch.zeros(216, dtype=torch.float32, device=dev)
        gs = [None if g is None else g.contiguous() for g in grads]      # named: the pointers must outlive the call
        if gs[0] is None:
            gs[0] = torch.zeros(co, n, k + 2, k + 2, dtype=torch.float32, device=dev)   # the interior fold owns the overwrite
        arr = (ctypes.c_void_p * 9)()
        for i in range(9):
            arr[i] = None if gs[i] is None else gs[i].data_ptr()
        _lib.check(_lib.lib().pn_pack_fold_set_backward(co, n, k, _lib.ptr(w2c), _lib.ptr(w3c), arr,
                                                        None if gs[9] is None else _lib.ptr(gs[9]), _lib.ptr(dw2), _lib.ptr(dw3),
                                                        _lib.current_stream()), "pn_pack_fold_set_backward")
        return dw2, dw3.view(8, 1, 3, 3, 3)


def _use_kernels(t):
    """CUDA tensors take the kernel path (the only one the product runs: PackNet01.forward refuses CPU tensors)."""
    return t.is_cuda


def fold_set(w2, w3):
    """-> (W_eff, Wtop, Wbottom, Wleft, Wright, Wtl, Wtr, Wbl, Wbr, S); kernels on CUDA tensors (frame weights
    channels-last), PyTorch ops otherwise (all OIHW; the CPU form exists for the algebra tests only: PackNet01.forward
    refuses CPU tensors)."""
    if _use_kernels(w2):
        return _FoldSetCUDA.apply(w2, w3)
    return fold_set_torch(w2, w3)


# ---------------------------------------------------------------------------------------------------------------
# frame terms (O(perimeter)): fp32 matmuls on the border rows / columns
# ---------------------------------------------------------------------------------------------------------------
def _row_conv(row, wt, m, k):
    """row [B, L, n] (one border row / column of xs, zero beyond both ends), wt [Co, n, m, k+2] ->
    [B, m, L, Co]: out[b, a, l, co] = sum_{n, e} wt[co, n, a, e] * row[b, l + e - (m+1), n]."""
    B, L, n = row.shape
    co = wt.shape[0]
    win = F.pad(row, (0, 0, m + 1, m + 1)).unfold(1, k + 2, 1)                          # [B, L, n, k+2]
    o = win.reshape(B * L, n * (k + 2)) @ wt.permute(1, 3, 2, 0).reshape(n * (k + 2), m * co)
    return o.view(B, L, m, co).permute(0, 2, 1, 3)


_const_cache = {}


def _class_consts(k, w, device):
    """(tap mask [2m+1, k]: which taps of a k-kernel stay inside the map for each border class; column class of
    every x in [0, w); the outer product of the tap mask with itself as a [(2m+1)^2, k*k] matrix)"""
    key = (k, w, str(device))
    if key not in _const_cache:
        m = k // 2
        g = 2 * m + 1
        t = torch.arange(g).view(g, 1)
        kk = torch.arange(k).view(1, k)
        mask = ((kk >= m - t) & (kk <= 3 * m - t)).to(torch.float32)
        cx = torch.cat([torch.arange(m), torch.full((w - 2 * m,), m, dtype=torch.long), torch.arange(m + 1, g)])
        mask2 = (mask.view(g, 1, k, 1) * mask.view(1, g, 1, k)).reshape(g * g, k * k)   # [(t,u), (ky,kx)]
        _const_cache[key] = (mask.to(device), cx.to(device), mask2.to(device))
    return _const_cache[key]


def bias_classes(S, b3, k, w, device):
    """Conv3d-bias term conv2d(b3 * 1_map, sum_c' W2): depends on a pixel only through its border class.
    -> (beta [Co]: the interior value, dB [2m+1, 2m+1, Co]: class value minus beta).  Two small matmuls:
    Sb[co, tap] = sum_f b3[f] S[co, f, tap];  bclass[(t,u), co] = sum_tap mask[t, ky] mask[u, kx] Sb[co, tap]."""
    m = k // 2
    g = 2 * m + 1
    co = S.shape[0]
    mask2 = _class_consts(k, w, device)[2].to(S.dtype)
    Sb = torch.matmul(b3, S.reshape(co, 8, k * k))                                    # [Co, k*k]
    bclass = torch.matmul(mask2, Sb.t()).view(g, g, co)                               # [m, m] = interior
    beta = bclass[m, m]
    return beta, bclass - beta


def frame_strips(top, bot, left, right, folds, b3, k):
    """The frame terms with PyTorch ops (the definition; CPU tests and GPU cross-check of pn_pack_frame_*).
    -> (beta [Co], top / bottom strips [B, m, w, Co], left / right strips [B, h, m, Co]): what must be ADDED to
    conv(xs, W_eff) + b2 + beta.  folds = fold_set_torch(w2, w3) (all OIHW)."""
    _, Wt, Wb, Wl, Wr, Wtl, Wtr, Wbl, Wbr, S = folds
    co = S.shape[0]
    m = k // 2
    B, w, _ = top.shape
    h = left.shape[1]
    beta, dB = bias_classes(S, b3, k, w, top.device)
    cx = _class_consts(k, w, top.device)[1]
    # ---- ring rows / columns; every index runs as (m-1-idx) from the border, hence the flips
    top_s = dB[0:m][:, cx].unsqueeze(0) - _row_conv(top, Wt, m, k).flip(1)              # [B, m, w, Co]
    bot_s = dB[m + 1:][:, cx].unsqueeze(0) - _row_conv(bot, Wb, m, k).flip(1)
    left_s = -_row_conv(left, Wl.transpose(2, 3), m, k).flip(1).permute(0, 2, 1, 3)     # [B, h, m, Co]
    right_s = -_row_conv(right, Wr.transpose(2, 3), m, k).flip(1).permute(0, 2, 1, 3)
    # bias delta of the side columns on the rows the top / bottom strips do not cover
    zpad = torch.zeros(m, m, co, dtype=S.dtype, device=S.device)
    left_s = left_s + torch.cat([zpad, dB[m, 0:m].unsqueeze(0).expand(h - 2 * m, m, co), zpad]).unsqueeze(0)
    right_s = right_s + torch.cat([zpad, dB[m, m + 1:].unsqueeze(0).expand(h - 2 * m, m, co), zpad]).unsqueeze(0)
    # ---- ring corners: counted by a row term and a column term, give one back
    def corner(px, wc):
        return torch.einsum("bn,onkl->bklo", px, wc).flip(1, 2)                        # [B, m, m, Co]
    top_s = top_s + F.pad(corner(top[:, 0], Wtl), (0, 0, 0, w - m)) + F.pad(corner(top[:, w - 1], Wtr), (0, 0, w - m, 0))
    bot_s = bot_s + F.pad(corner(bot[:, 0], Wbl), (0, 0, 0, w - m)) + F.pad(corner(bot[:, w - 1], Wbr), (0, 0, w - m, 0))
    return beta, top_s, bot_s, left_s, right_s


def frame_term_specs(h, w, n, k):
    """The eight frame terms in the vocabulary of pn_pack_frame_* (include/packnet_b200.h): which border line and folded
    weight a term reads, the strides of the channels-last weight, and the affine map (a, l) -> (row, col) into z.
    Every index counts from the border inwards as (m-1-idx) -- the flips of frame_strips()."""
    m = k // 2
    side = dict(A=m, A2=1, KE=k + 2, pad=m + 1, alpha=-1.0, px=0)
    rowt = dict(side, L=w, w_sco=m * (k + 2) * n, w_sa=(k + 2) * n, w_se=n, bias_mode=1)      # weight [Co][m][k+2][n]
    colt = dict(side, L=h, w_sco=(k + 2) * m * n, w_sa=n, w_se=m * n, bias_mode=2)            # weight [Co][k+2][m][n]
    corn = dict(A=m * m, A2=m, KE=1, pad=0, L=1, alpha=1.0, w_sco=m * m * n, w_sa=n, w_se=0, bias_mode=0)
    z4 = dict(ra1=0, ra2=0, rl=0, ca1=0, ca2=0, cl=0)
    return [
        dict({**z4, **rowt}, name="top", line="top", r0=m - 1, ra1=-1, c0=0, cl=1),
        dict({**z4, **rowt}, name="bottom", line="bottom", r0=h - 1, ra1=-1, c0=0, cl=1),
        dict({**z4, **colt}, name="left", line="left", r0=0, rl=1, c0=m - 1, ca1=-1),
        dict({**z4, **colt}, name="right", line="right", r0=0, rl=1, c0=w - 1, ca1=-1),
        dict({**z4, **corn}, name="tl", line="top", px=0, r0=m - 1, ra1=-1, c0=m - 1, ca2=-1),
        dict({**z4, **corn}, name="tr", line="top", px=w - 1, r0=m - 1, ra1=-1, c0=w - 1, ca2=-1),
        dict({**z4, **corn}, name="bl", line="bottom", px=0, r0=h - 1, ra1=-1, c0=m - 1, ca2=-1),
        dict({**z4, **corn}, name="br", line="bottom", px=w - 1, r0=h - 1, ra1=-1, c0=w - 1, ca2=-1),
    ]


def _frame_flags():
    """A/B switch of the frame GEMMs' K split (PN_FRAME_KSPLIT=0 -> PN_FRAME_FLAG_NO_KSPLIT: the round-1 schedule)"""
    import os
    return 2 if os.environ.get("PN_FRAME_KSPLIT") == "0" else 0


class _FrameApplyCUDA(torch.autograd.Function):
    """z += frame terms, in place (one launch); backward: border-line, folded-weight and bias-class gradients
    (three launches).  weights = the eight channels-last frame folds in FOLD_ORDER[1:]."""

    _templates = {}

    @staticmethod
    def _desc(z_shape, n, k, lines, weights, dlines=None, dws=None):
        """pn_frame_desc for this call: the shape-dependent part (term table, strides, maps) is built once per shape and
        copied; only the pointers change from call to call."""
        from ._lib_conv import FrameDesc
        key = (tuple(z_shape), n, k)
        tpl = _FrameApplyCUDA._templates.get(key)
        if tpl is None:
            B, h, w, co = z_shape
            d = FrameDesc()
            d.batch, d.height, d.width, d.cout, d.n, d.ksize = B, h, w, co, n, k
            specs = frame_term_specs(h, w, n, k)
            d.num_terms = len(specs)
            where = []
            for i, sp in enumerate(specs):
                t = d.terms[i]
                lfull = w if sp["line"] in ("top", "bottom") else h
                t.line_bstride = t.dline_bstride = lfull * n
                for f in ("w_sco", "w_sa", "w_se", "L", "A", "A2", "KE", "pad", "r0", "ra1", "ra2", "rl", "c0", "ca1", "ca2", "cl",
                          "alpha", "bias_mode"):
                    setattr(t, f, sp[f])
                where.append((sp["line"], sp["px"] * n * 4, sp["name"]))
            tpl = (d, where)
            _FrameApplyCUDA._templates[key] = tpl
        d = FrameDesc.from_buffer_copy(tpl[0])
        for i, (line, off, name) in enumerate(tpl[1]):
            t = d.terms[i]
            t.line = lines[line].data_ptr() + off
            t.w = weights[name].data_ptr()
            if dlines is not None:
                t.dline = dlines[line].data_ptr() + off
                t.dw = dws[name].data_ptr()
        return d

    @staticmethod
    def forward(ctx, z, top, bot, left, right, Wt, Wb, Wl, Wr, Wtl, Wtr, Wbl, Wbr, dB, k):
        from . import _lib
        _lib.require_f32(z, top, dB)
        n = top.shape[2]
        lines = {"top": top.contiguous(), "bottom": bot.contiguous(), "left": left.contiguous(), "right": right.contiguous()}
        weights = dict(zip(FOLD_ORDER[1:], (t.contiguous() for t in (Wt, Wb, Wl, Wr, Wtl, Wtr, Wbl, Wbr))))
        dBc = dB.detach().contiguous()
        d = _FrameApplyCUDA._desc(tuple(z.shape), n, k, lines, weights)
        d.flags = _frame_flags()
        _lib.check(_lib.lib().pn_pack_frame_forward(ctypes.byref(d), _lib.ptr(dBc), _lib.ptr(z), _lib.current_stream()),
                   "pn_pack_frame_forward")
        ctx.mark_dirty(z)
        ctx.save_for_backward(*lines.values(), *weights.values())
        ctx.k, ctx.n, ctx.dB_shape = k, n, tuple(dB.shape)
        return z

    @staticmethod
    def backward(ctx, gz):
        from . import _lib
        saved = ctx.saved_tensors
        lines = dict(zip(("top", "bottom", "left", "right"), saved[:4]))
        weights = dict(zip(FOLD_ORDER[1:], saved[4:]))
        gz = gz.contiguous()
        B, h, w, co = gz.shape
        n, k = ctx.n, ctx.k
        flat = torch.zeros(B * (2 * w + 2 * h) * n, dtype=torch.float32, device=gz.device)     # one memset for the four lines
        o1, o2, o3 = B * w * n, 2 * B * w * n, 2 * B * w * n + B * h * n
        dlines = {"top": flat[:o1].view(B, w, n), "bottom": flat[o1:o2].view(B, w, n),
                  "left": flat[o2:o3].view(B, h, n), "right": flat[o3:].view(B, h, n)}
        # the eight weight gradients in ONE zeroed buffer: the library splits their GEMMs over the samples (atomic adds)
        sizes = [(t.numel() + 3) // 4 * 4 for t in weights.values()]
        wflat = torch.zeros(sum(sizes), dtype=torch.float32, device=gz.device)
        dws, o = {}, 0
        for (name, t), sz in zip(weights.items(), sizes):
            dws[name] = wflat[o:o + t.numel()].view(t.shape)
            o += sz
        gdB = torch.zeros(ctx.dB_shape, dtype=torch.float32, device=gz.device)
        d = _FrameApplyCUDA._desc((B, h, w, co), n, k, lines, weights, dlines, dws)
        d.flags = 1 | _frame_flags()      # PN_FRAME_FLAG_DW_ZEROED
        _lib.check(_lib.lib().pn_pack_frame_backward(ctypes.byref(d), _lib.ptr(gz), _lib.ptr(gdB), _lib.current_stream()),
                   "pn_pack_frame_backward")
        return (gz, dlines["top"], dlines["bottom"], dlines["left"], dlines["right"]) + tuple(dws[nm] for nm in FOLD_ORDER[1:]) \
            + (gdB, None)


def prefold(w2, w3, b3, w_packed):
    """The weight-only half of pack_conv_folded -- the nine folds, S and the Conv3d-bias classes -- for a packed map that is
    `w_packed` pixels wide.  It depends on parameters alone, so networks.PackNet01 issues it on a side stream at the start
    of the forward (0.31 ms of small launches per step, and through autograd's stream replay the 0.50 ms of fold backward
    launches, next to the convolutions instead of between them).  -> (folds, beta, dB)"""
    folds = fold_set(w2, w3)
    beta, dB = bias_classes(folds[9], b3, w2.shape[2], w_packed, w2.device)
    return folds, beta, dB


def pack_conv_folded(x, w2, b2, w3, b3, conv, pre=None):
    """z = Conv2d(W2, b2)(pad(Conv3d(W3, b3)(packing(x)))) on NHWC maps: x [B,2h,2w,C] -> z [B,h,w,Co].
    `conv(xs, weight, bias)` is the O(area) convolution (functional.conv2d on the GPU); `pre` = prefold(...) when the caller
    computed the weight-only half ahead of time."""
    co, c8, k, _ = w2.shape
    m = k // 2
    B, H, W, C = x.shape
    h, w = H // 2, W // 2
    if c8 != 32 * C or H % 2 or W % 2:
        raise ValueError("pack_conv_folded: weight %s does not match input %s" % (tuple(w2.shape), tuple(x.shape)))
    if k not in (3, 5):
        raise ValueError("pack_conv_folded: kernel size %d (the fold kernels cover 3 and 5)" % k)
    if h < 2 * m + 1 or w < 2 * m + 1:
        raise ValueError("pack_conv_folded: packed map %dx%d smaller than the frame of a %dx%d kernel" % (h, w, k, k))
    xs, top, bot, left, right = space_to_depth_borders(x.contiguous())
    if pre is not None and _use_kernels(x):
        folds, beta, dB = pre
        z = conv(xs, folds[0], b2 + beta)
        return _FrameApplyCUDA.apply(z, top, bot, left, right, *folds[1:9], dB, k)
    folds = fold_set(w2, w3)
    if _use_kernels(x):
        beta, dB = bias_classes(folds[9], b3, k, w, x.device)
        z = conv(xs, folds[0], b2 + beta)
        return _FrameApplyCUDA.apply(z, top, bot, left, right, *folds[1:9], dB, k)
    beta, top_s, bot_s, left_s, right_s = frame_strips(top, bot, left, right, folds, b3, k)
    z = conv(xs, folds[0], b2 + beta)
    z[:, :m] += top_s
    z[:, h - m:] += bot_s
    z[:, :, :m] += left_s
    z[:, :, w - m:] += right_s
    return z
