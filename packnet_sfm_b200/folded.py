"""Folded pack block: the two linear stages of PackLayerConv3d (layers01.py:239-247) as ONE convolution.

    reference:  xs = packing(x)                      [B, n=4C, h, w]        (layers01.py:126-148, 243)
                T  = Conv3d(1->8, 3x3x3, pad 1)(xs)  [B, 8n,  h, w]         (layers01.py:236-237, 244-246)
                z  = Conv2d(8n -> Co, k x k)(zero-pad k//2 (T))             (layers01.py:28-30, 36, 247)

Nothing but the zero padding sits between the Conv3d and the Conv2d, so z is linear in xs and, away from the border,
equals a single Conv2d(n -> Co, (k+2) x (k+2)) of xs whose weight is the FULL convolution of the two kernels:

    W_eff[co, c'', ey, ex] = sum_{f, dc, dy, dx}  W2[co, f*n + (c''-dc+1), ey-dy, ex-dx] * W3[f, dc, dy, dx]

(= conv_transpose3d of W2 viewed as [Co, 8, n, k, k] with W3; depth index c'' is clipped to [0, n) exactly as the
Conv3d's zero padding in depth does).  The reduction length drops from 8n*k*k to n*(k+2)^2 -- pack1: 51 200 -> 12 544
(4.08x fewer MACs), pack2..5: 288C -> 100C (2.88x) -- and the [B, 8n, h, w] intermediate (1.0 GB for pack1 at B=4,
192x640), its bf16 split and the feature-stencil kernels disappear from the step.

What the single convolution gets wrong is the FRAME of width m = k//2: the reference zero-pads T, while the folded
kernel implicitly continues the Conv3d one pixel outside the map (the ring q at distance 1, where T_full(q) != 0 because
its 3x3 footprint still touches the border row / column of xs) and keeps the Conv3d bias there.  Both are repaired
exactly with thin strips:

    z = conv(xs, W_eff) + b2 + beta          (beta = interior value of the Conv3d-bias term)
        - top - bottom - left - right        (ring row / column folded into 1 x (k+2) resp. (k+2) x 1 kernels that
                                              read only the border row / column of xs)
        + 4 corner blocks                    (the ring corners are in a row AND a column term: inclusion-exclusion)
        + dB                                 (bias term on the frame: fewer taps of W2 see an in-map T)

All the strip work is O(perimeter) and is expressed here with PyTorch ops (weights folded in fp32 with TF32 off, forward
and backward); the O(area) convolution runs on the tcgen05 engine (functional.conv2d).  The s2d tensor uses the channel
order (i, j, c) -- 2C contiguous floats of the NHWC source per half-row -- instead of the reference's (c, i, j); the folded
weights are permuted accordingly, so no tensor in the reference's channel order is ever materialised.

Status (round 1): algebra verified on the CPU against the reference composition in float64 (tests/test_folded_cpu.py:
values and all gradients to 1e-12); NOT yet run on the B200 -- PackLayerConv3d uses it only when
functional.set_pack_fold(True) / PN_PACK_FOLD=1 is given."""
import torch
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------------------------
# exact-fp32 region: cuDNN may not use TF32 for the weight folds, neither in the forward nor in the backward
# ---------------------------------------------------------------------------------------------------------------
class _ExactFP32(torch.autograd.Function):
    """outs = fn(*args) with TF32 disabled; the backward re-runs fn under the same flags (autograd would otherwise
    execute the recorded graph outside the context, with PyTorch's default cudnn.allow_tf32=True)."""

    @staticmethod
    def forward(ctx, fn, *args):
        ctx.fn = fn
        ctx.save_for_backward(*args)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            outs = fn(*[a.detach() for a in args])
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        need = ctx.needs_input_grad[1:]
        args = [a.detach().requires_grad_(n) for a, n in zip(ctx.saved_tensors, need)]
        with torch.enable_grad(), torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            outs = ctx.fn(*args)
            pairs = [(o, g) for o, g in zip(outs, gouts) if g is not None and o.requires_grad]
            wrt = [a for a, n in zip(args, need) if n]
            grads = torch.autograd.grad([o for o, _ in pairs], wrt, [g for _, g in pairs], allow_unused=True) if pairs and wrt else []
        it = iter(grads)
        return (None,) + tuple(next(it) if n else None for n in need)


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
            gx6 = g_xs.view(B, h, w, 2, 2, C).permute(0, 1, 3, 2, 4, 5).contiguous()     # [B,h,i,w,j,C], fresh tensor
        if g_top is not None:
            gx6[:, 0] += g_top.view(B, w, 2, 2, C).permute(0, 2, 1, 3, 4)
        if g_bot is not None:
            gx6[:, h - 1] += g_bot.view(B, w, 2, 2, C).permute(0, 2, 1, 3, 4)
        if g_left is not None:
            gx6[:, :, :, 0] += g_left.view(B, h, 2, 2, C)
        if g_right is not None:
            gx6[:, :, :, w - 1] += g_right.view(B, h, 2, 2, C)
        return gx6.view(B, 2 * h, 2 * w, C)


def space_to_depth_borders(x):
    return _SpaceToDepthBorders.apply(x)


# ---------------------------------------------------------------------------------------------------------------
# weight folds
# ---------------------------------------------------------------------------------------------------------------
def _ct3(w2s, w3s):
    """full convolution of a W2 slab [Co, 8, n, a, b] with a W3 slab [8, 1, 3, c, d] -> [Co, n, a+c-1, b+d-1]
    (depth clipped to [0, n): the Conv3d pads the depth with zeros, those taps never meet data)."""
    return F.conv_transpose3d(w2s.contiguous(), w3s.contiguous(), padding=(1, 0, 0)).squeeze(1)


def _perm_n(t):
    """folded weight [Co, n=(c,i,j), ...] in the reference's s2d channel order -> ours, (i,j,c)."""
    co, n = t.shape[:2]
    return t.reshape(co, n // 4, 4, *t.shape[2:]).transpose(1, 2).reshape(co, n, *t.shape[2:])


def _row_conv(row, wt, m, k):
    """row [B, L, n] (one border row / column of xs, zero beyond both ends), wt [Co, n, m, k+2] ->
    [B, m, L, Co]: out[b, a, l, co] = sum_{n, e} wt[co, n, a, e] * row[b, l + e - (m+1), n]."""
    B, L, n = row.shape
    co = wt.shape[0]
    wgt = wt.permute(2, 0, 1, 3).reshape(m * co, n, 1, k + 2)
    o = F.conv2d(row.permute(0, 2, 1).unsqueeze(2), wgt, padding=(0, m + 1))          # [B, m*Co, 1, L]
    return o.view(B, m, co, L).permute(0, 1, 3, 2)


def _fold_all(top, bot, left, right, w2, b2, w3, b3):
    """-> (W_eff [Co, n, k+2, k+2] in (i,j,c) channel order, bias [Co], top / bottom strips [B, m, w, Co],
    left / right strips [B, h, m, Co]) -- the strips are what must be ADDED to conv(xs, W_eff) + bias."""
    co, c8, k, _ = w2.shape
    n = c8 // 8
    m = k // 2
    B, w, _ = top.shape
    h = left.shape[1]
    w2r = w2.reshape(co, 8, n, k, k)
    w_eff = _perm_n(_ct3(w2r, w3))
    # ---- Conv3d-bias term: conv2d(b3 * 1_map, sum_c' W2) depends on the pixel only through its border class
    g = 2 * m + 1
    S = w2r.sum(2)                                                                     # [Co, 8, k, k]
    bclass = F.conv2d(b3.view(1, 8, 1, 1).expand(1, 8, g, g), S, padding=m)[0]         # [Co, g, g]; [m, m] = interior
    beta = bclass[:, m, m]
    dB = (bclass - beta.view(co, 1, 1)).permute(1, 2, 0)                               # [g, g, Co]
    cx = torch.cat([torch.arange(m), torch.full((w - 2 * m,), m, dtype=torch.long), torch.arange(m + 1, g)]).to(w2.device)
    lo, hi = slice(0, m), slice(m + 1, k)
    # ---- ring rows / columns (weights: ky or kx restricted to the taps that reach the ring, W3 to the face that
    #      still touches the map); every index runs as (m-1-idx) from the border, hence the flips
    Wt = _perm_n(_ct3(w2r[:, :, :, lo, :], w3[:, :, :, 2:3, :]))                        # [Co, n, m, k+2]
    Wb = _perm_n(_ct3(w2r[:, :, :, hi, :], w3[:, :, :, 0:1, :]))
    Wl = _perm_n(_ct3(w2r[:, :, :, :, lo], w3[:, :, :, :, 2:3])).transpose(2, 3)        # [Co, n, m, k+2]
    Wr = _perm_n(_ct3(w2r[:, :, :, :, hi], w3[:, :, :, :, 0:1])).transpose(2, 3)
    top_s = dB[0:m][:, cx].unsqueeze(0) - _row_conv(top, Wt, m, k).flip(1)              # [B, m, w, Co]
    bot_s = dB[m + 1:][:, cx].unsqueeze(0) - _row_conv(bot, Wb, m, k).flip(1)
    left_s = -_row_conv(left, Wl, m, k).flip(1).permute(0, 2, 1, 3)                     # [B, h, m, Co]
    right_s = -_row_conv(right, Wr, m, k).flip(1).permute(0, 2, 1, 3)
    # bias delta of the side columns on the rows the top / bottom strips do not cover
    zpad = torch.zeros(m, m, co, dtype=w2.dtype, device=w2.device)
    left_s = left_s + torch.cat([zpad, dB[m, 0:m].unsqueeze(0).expand(h - 2 * m, m, co), zpad]).unsqueeze(0)
    right_s = right_s + torch.cat([zpad, dB[m, m + 1:].unsqueeze(0).expand(h - 2 * m, m, co), zpad]).unsqueeze(0)
    # ---- ring corners: counted by a row term and a column term, give one back
    def corner(px, ys, xs_, dy, dx):
        wc = _perm_n(_ct3(w2r[:, :, :, ys, xs_], w3[:, :, :, dy:dy + 1, dx:dx + 1]))    # [Co, n, m, m]
        return torch.einsum("bn,onkl->bklo", px, wc).flip(1, 2)                        # [B, m, m, Co]
    tl = corner(top[:, 0], lo, lo, 2, 2)
    tr = corner(top[:, w - 1], lo, hi, 2, 0)
    bl = corner(bot[:, 0], hi, lo, 0, 2)
    br = corner(bot[:, w - 1], hi, hi, 0, 0)
    top_s = top_s + F.pad(tl, (0, 0, 0, w - m)) + F.pad(tr, (0, 0, w - m, 0))
    bot_s = bot_s + F.pad(bl, (0, 0, 0, w - m)) + F.pad(br, (0, 0, w - m, 0))
    return w_eff, b2 + beta, top_s, bot_s, left_s, right_s


def pack_conv_folded(x, w2, b2, w3, b3, conv):
    """z = Conv2d(W2, b2)(pad(Conv3d(W3, b3)(packing(x)))) on NHWC maps: x [B,2h,2w,C] -> z [B,h,w,Co].
    `conv(xs, weight, bias)` is the O(area) convolution (functional.conv2d on the GPU)."""
    co, c8, k, _ = w2.shape
    m = k // 2
    B, H, W, C = x.shape
    h, w = H // 2, W // 2
    if c8 != 32 * C or H % 2 or W % 2:
        raise ValueError("pack_conv_folded: weight %s does not match input %s" % (tuple(w2.shape), tuple(x.shape)))
    if h < 2 * m + 1 or w < 2 * m + 1:
        raise ValueError("pack_conv_folded: packed map %dx%d smaller than the frame of a %dx%d kernel" % (h, w, k, k))
    xs, top, bot, left, right = space_to_depth_borders(x.contiguous())
    w_eff, bias, top_s, bot_s, left_s, right_s = _ExactFP32.apply(_fold_all, top, bot, left, right, w2, b2, w3, b3)
    z = conv(xs, w_eff, bias)
    z[:, :m] += top_s
    z[:, h - m:] += bot_s
    z[:, :, :m] += left_s
    z[:, :, w - m:] += right_s
    return z
