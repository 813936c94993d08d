"""Autograd functions over the C-ABI layer kernels.  Every tensor here is an NHWC feature map stored as a
plain contiguous [B,H,W,C] CUDA tensor (the same bytes as a torch.channels_last NCHW tensor).

  conv2d            tcgen05 implicit-GEMM convolution (+bias); backward = data gradient through the same
                    kernel on the flipped/transposed packed weight, weight gradient through the wgrad kernel
  groupnorm_elu     GroupNorm(16)+ELU (optionally of x + x2), output optionally into a channel window
  pack_features     space-to-depth + Conv3d(1->8) feature stencil   (PackLayerConv3d, layers01.py:239-245)
  unpack_features   Conv3d(1->8) feature stencil + depth-to-space   (UnpackLayerConv3d, layers01.py:281-285)

Precision of the tensor-core GEMMs (include/packnet_b200.h): PRECISION_BF16X3 (default: error-compensated bf16
split, 16 mantissa bits, meets the 1e-3 depth parity bar at twice the tf32 MMA rate), PRECISION_TF32X3 (22 bits),
PRECISION_TF32X1 (what cuDNN gives the reference on Ampere+ with PyTorch defaults; fails the parity bar)."""
import contextlib
import ctypes
import os
import weakref

import torch

from . import _lib
from ._lib_conv import ConvDesc, PRECISION_TF32X1, PRECISION_BF16X1, PRECISION_TF32X3, PRECISION_BF16X3, MODE_AUTO

# Defaults since round 2 = what the B200 measured faster (gpurun r02a, B=4 192x640, ms per step): pack fold 37.9 -> 30.9, im2col
# first layer -0.2, tiled weight-gradient unpack -0.2; the tiled weight PACK was not faster (1.85 vs 1.90 ms) and stays off.
# PN_<NAME>=0 / =1 in the environment overrides a default for A/B runs.
def _env(name, default):
    v = os.environ.get(name)
    return default if v is None else v == "1"


_state = {"precision": PRECISION_BF16X3, "mode": MODE_AUTO, "pack_fold": _env("PN_PACK_FOLD", True),
          "pack_fold_min_pixels": int(os.environ.get("PN_PACK_FOLD_MIN_PIXELS", "1920")),
          "im2col_first": _env("PN_IM2COL_FIRST", True),
          "unpack_tiled": _env("PN_UNPACK_TILED", True),
          "pack_tiled": _env("PN_PACK_TILED", False),
          # GroupNorm+ELU kernels write the bf16 operand pair of their output for the convolution that consumes it
          "gn_emit_split": _env("PN_GN_EMIT_SPLIT", True),
          # weight gradients of stored weights on a side stream during the backward (joined by an autograd end-of-backward callback)
          "wgrad_stream": _env("PN_WGRAD_STREAM", True)}


def set_pack_tiled(on):
    """Off by default (measured on the B200: not faster than the gather): per-step weight packing of the bf16
    precisions through shared memory (pn_conv2d_pack_weight_tiled) instead of the strided element-per-thread gather."""
    prev, _state["pack_tiled"] = _state["pack_tiled"], bool(on)
    return prev


def set_unpack_tiled(on):
    """On by default since round 2: weight-gradient re-layout [Cout][tap][Cin] -> OIHW through shared memory
    (pn_conv2d_unpack_weight_grad_tiled) instead of the strided element-per-thread gather."""
    prev, _state["unpack_tiled"] = _state["unpack_tiled"], bool(on)
    return prev


def set_im2col_first(on):
    """On by default since round 2: evaluate the network's first convolution (3 -> 64, 5x5) as a 1x1 convolution
    over its im2col tensor (conv2d_im2col) instead of 25 tap items that each fill 3 of 64 reduction lanes."""
    prev, _state["im2col_first"] = _state["im2col_first"], bool(on)
    return prev


def im2col_first_enabled():
    return _state["im2col_first"]


def set_pack_fold(on, min_pixels=None):
    """Pack layers as ONE folded convolution of the space-to-depth tensor (packnet_sfm_b200/folded.py) instead of
    feature stencil + convolution over the 8x-inflated channel count.  ON by default since round 2 (B200: 37.9 -> 30.9 ms/step).
    min_pixels: fold only layers whose packed map has at least that many pixels -- on the small maps (pack4: 12x40,
    pack5: 6x20 at 192x640) the frame is 20-40 % of the map and the layer is bound by streaming its weights, which the
    fold has to read once more (default 1920 = pack1..pack3 at 192x640)."""
    _state["pack_fold"] = bool(on)
    if min_pixels is not None:
        _state["pack_fold_min_pixels"] = int(min_pixels)


def pack_fold_enabled(packed_pixels=None):
    return _state["pack_fold"] and (packed_pixels is None or packed_pixels >= _state["pack_fold_min_pixels"])


_wgrad_side = {}          # device index -> {"stream": side stream, "dirty": launches since the last join}


def _wgrad_side_stream(cur):
    """The side stream of the weight-gradient launches on `cur`'s device.  Every use inside a backward pass queues an
    end-of-backward callback; the first one to run makes the stream the backward ran on wait for the side stream, so that
    `loss.backward()` returns with every gradient ordered before whatever the caller enqueues next (optimizer, all-reduce, a
    read of .grad).  (One callback per launch rather than one per pass: a pass that died with an exception must not leave a
    stale "already queued" mark behind.)"""
    st = _wgrad_side.get(cur.device.index)
    if st is None:
        with torch.cuda.device(cur.device):
            st = _wgrad_side[cur.device.index] = {"stream": torch.cuda.Stream(), "dirty": False}
    side = st["stream"]
    st["dirty"] = True

    def join():
        if st["dirty"]:
            st["dirty"] = False
            cur.wait_stream(side)

    torch.autograd.Variable._execution_engine.queue_callback(join)
    return side


def wait_for_wgrad_stream(*tensors):
    """For the consumer of a weight gradient produced with wgrad_side=True: order the current stream behind the side stream
    and tell the allocator that `tensors` (allocated on the side stream) are used here."""
    if not torch.cuda.is_available():
        return
    cur = torch.cuda.current_stream()
    st = _wgrad_side.get(cur.device.index)
    if st is None:
        return
    cur.wait_stream(st["stream"])
    for t in tensors:
        if t is not None and t.is_cuda:
            t.record_stream(cur)


def set_wgrad_stream(on):
    prev, _state["wgrad_stream"] = _state["wgrad_stream"], bool(on)
    return prev


def prefold_stream_enabled():
    """Weight folds of the folded pack layers on a side stream at the start of PackNet01.forward (PN_PREFOLD_STREAM=0: inline,
    on the main stream, where the layer runs)."""
    return _env("PN_PREFOLD_STREAM", True)


def set_precision(p):
    assert p in (PRECISION_TF32X1, PRECISION_BF16X1, PRECISION_TF32X3, PRECISION_BF16X3)
    _state["precision"] = p


def is_bf16(p):
    return p in (PRECISION_BF16X1, PRECISION_BF16X3)


def is_split(p):
    return p in (PRECISION_TF32X3, PRECISION_BF16X3)


def channel_align(p=None):
    """Channel multiple an NHWC operand needs (16-byte TMA row pitch): 4 fp32 or 8 bf16 elements."""
    return 8 if is_bf16(_state["precision"] if p is None else p) else 4


def get_precision():
    return _state["precision"]


def set_conv_mode(m):
    _state["mode"] = m


def _stream():
    return _lib.current_stream()


_errflag = {}


def error_flag():
    """Pinned host word the conv kernels write 0xDEAD00xx into when a pipeline wait times out (they trap
    instead of hanging); being host memory it stays readable after the CUDA context died."""
    if "t" not in _errflag:
        _errflag["t"] = torch.zeros(4, dtype=torch.int32).pin_memory()
    return _errflag["t"]


def read_error_flag():
    return int(error_flag()[0].item()) & 0xFFFFFFFF


def _p(t):
    return _lib.ptr(t) if t is not None else None


def _residual(x):
    lo = torch.empty_like(x)
    _lib.check(_lib.lib().pn_tf32_residual(_lib.ptr(x), _lib.ptr(lo), x.numel(), _stream()), "pn_tf32_residual")
    return lo


def _attach_split(t, hi, lo):
    """remember the bf16 operand pair the producing kernel wrote next to a tensor (valid while the tensor is unmodified)"""
    t._pn_split = (hi, lo, t._version)


def _operands(x, precision):
    """fp32 NHWC tensor -> (hi, lo) tensor-core operands of the given precision (lo is None for the X1 modes).  A tensor whose
    producer (GroupNorm+ELU forward / backward) already emitted its bf16 pair carries it along: no split launch."""
    if precision == PRECISION_BF16X3:
        pair = getattr(x, "_pn_split", None)
        if pair is not None and pair[2] == x._version and pair[0].shape == x.shape:
            return pair[0], pair[1]
    if is_bf16(precision):
        hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
        _lib.check(_lib.lib().pn_split_bf16(_lib.ptr(x), _lib.ptr(hi), _lib.ptr(lo), x.numel(), _stream()), "pn_split_bf16")
        return hi, (lo if is_split(precision) else None)
    return x, (_residual(x) if is_split(precision) else None)


def _pack_weight(w, transposed, precision):
    cout, cin, k, _ = w.shape
    n = ctypes.c_size_t(0)
    lib = _lib.lib()
    _lib.check(lib.pn_conv2d_packed_weight_elems(cout, cin, k, int(transposed), precision, ctypes.byref(n)), "packed_weight_elems")
    dt = torch.bfloat16 if is_bf16(precision) else torch.float32
    wp = torch.empty(int(n.value), dtype=dt, device=w.device)
    lo = torch.empty_like(wp) if is_split(precision) else None
    if _state["pack_tiled"] and is_bf16(precision):
        kred = cout if transposed else cin
        rows_pad = int(n.value) // (((kred + 63) // 64) * k * k * 64)
        _lib.check(lib.pn_conv2d_pack_weight_tiled(_lib.ptr(w), _lib.ptr(wp), _p(lo), cout, cin, k, int(transposed), rows_pad, _stream()),
                   "pn_conv2d_pack_weight_tiled")
        return wp, lo
    _lib.check(lib.pn_conv2d_pack_weight(_lib.ptr(w), _lib.ptr(wp), _p(lo), cout, cin, k, int(transposed), precision, _stream()),
               "pn_conv2d_pack_weight")
    return wp, lo


def _conv_raw(x, x_lo, wp, wp_lo, bias, cout, ksize, precision):
    B, H, W, Cin = x.shape
    y = torch.empty(B, H, W, cout, dtype=torch.float32, device=x.device)
    d = ConvDesc(B, H, W, Cin, cout, ksize, precision, _state["mode"], 0)
    _lib.check(_lib.lib().pn_conv2d_forward(ctypes.byref(d), _lib.ptr(x), _p(x_lo), _lib.ptr(wp), _p(wp_lo), _p(bias),
                                            _lib.ptr(y), _lib.ptr(error_flag()), _stream()), "pn_conv2d_forward")
    return y


def _pad_channels(w, cin_tensor):
    """weight [Cout,Cin,k,k] whose Cin is smaller than the (4-aligned) channel count of the activation tensor."""
    if w.shape[1] == cin_tensor:
        return w
    pad = torch.zeros(w.shape[0], cin_tensor - w.shape[1], w.shape[2], w.shape[3], dtype=w.dtype, device=w.device)
    return torch.cat([w, pad], 1).contiguous()


def _conv_dgrad(g_hi, g_lo, wp, wp_lo, B, H, W, cin, cout, ksize, precision):
    """gx [B,H,W,cin] from the layer's FORWARD tiles (pn_conv2d_dgrad: MN-major reads, bf16 precisions)."""
    gx = torch.empty(B, H, W, cin, dtype=torch.float32, device=g_hi.device)
    d = ConvDesc(B, H, W, cin, cout, ksize, precision, _state["mode"], 0)
    _lib.check(_lib.lib().pn_conv2d_dgrad(ctypes.byref(d), _lib.ptr(g_hi), _p(g_lo), _lib.ptr(wp), _p(wp_lo), _lib.ptr(gx),
                                          _lib.ptr(error_flag()), _stream()), "pn_conv2d_dgrad")
    return gx


def _stored_weight(weight, cin_tensor, precision):
    """The NativeWeight of a parameter that packnet_sfm_b200.optim stores in the engine's layout, with current tiles -- or None
    (plain OIHW tensor: pack per call)."""
    nat = getattr(weight, "_pn_native", None)
    if nat is None or precision != PRECISION_BF16X3 or not nat.accepts(cin_tensor):
        return None
    if weight._version != nat.version:      # written through PyTorch since the tiles were made (load_state_dict, init)
        nat.owner.repack()
    return nat


class _Conv2d(torch.autograd.Function):
    """y = conv2d(x, weight, stride 1, zero pad k//2) + bias on NHWC tensors (nn.Conv2d + ConstantPad2d,
    layers01.py:28-30,36).

    Two weight paths.  (1) A parameter stored by optim.FlatAdam in the engine's layout: the forward and the data gradient
    read the bf16 tiles the optimizer step wrote, the weight gradient is accumulated straight into the flat gradient buffer
    (returned gradient: None -- `weight.grad` is a view of that buffer).  (2) Any other OIHW tensor (folded effective weights,
    tests, a model without the flat optimizer): packed per call, gradient returned in OIHW.  In the bf16 precisions the data
    gradient reads the FORWARD tiles in both paths (pn_conv2d_dgrad), so the transposed packing only exists for tf32."""

    @staticmethod
    def forward(ctx, x, weight, bias, wgrad_side=False):
        _lib.require_f32(x, weight, bias)
        x = x.contiguous()
        precision = _state["precision"]
        cout, cin_w, k, _ = weight.shape
        ctx.wgrad_side = bool(wgrad_side)
        nat = _stored_weight(weight, x.shape[3], precision)
        if nat is not None:
            wp, wp_lo = nat.hi, nat.lo
        else:
            w_eff = _pad_channels(weight.detach().contiguous(), x.shape[3])
            wp, wp_lo = _pack_weight(w_eff, False, precision)
        x_hi, x_lo = _operands(x, precision)
        y = _conv_raw(x_hi, x_lo, wp, wp_lo, bias.detach().contiguous() if bias is not None else None, cout, k, precision)
        # the weight gradient needs exactly the operand pair the forward used
        ctx.save_for_backward(x_hi, x_lo, weight)
        ctx.packed = (wp, wp_lo) if is_bf16(precision) else None
        ctx.nat = nat
        ctx.has_bias = bias is not None
        ctx.precision = precision
        return y

    @staticmethod
    def backward(ctx, gy):
        x_hi, x_lo, weight = ctx.saved_tensors
        precision = ctx.precision
        lib = _lib.lib()
        gy = gy.contiguous()
        B, H, W, Cin = x_hi.shape
        cout, cin_w, k, _ = weight.shape
        gx = gw = gb = None
        if ctx.needs_input_grad[0] or ctx.needs_input_grad[1]:
            # the output gradient has `cout` channels; bf16 operands need a multiple of 8
            g_hi, g_lo = _operands(gy, precision)
        side = None
        if ctx.needs_input_grad[1] and (ctx.nat is not None or ctx.wgrad_side) and _state["wgrad_stream"] and gy.is_cuda:
            # The weight gradient of a stored weight has no consumer inside the backward (it lands in the flat gradient buffer):
            # it runs on a side stream, next to the data-gradient / GroupNorm / stencil chain that IS the critical path -- its
            # CTAs fill the SMs the small-map launches leave idle and share SMs with the kernels that use no shared memory.
            cur = torch.cuda.current_stream()
            side = _wgrad_side_stream(cur)
            ready = torch.cuda.Event()
            ready.record(cur)                 # both operand pairs exist at this point of the main stream
            side.wait_event(ready)
        if ctx.needs_input_grad[0]:
            if ctx.packed is not None:
                gx = _conv_dgrad(g_hi, g_lo, ctx.packed[0], ctx.packed[1], B, H, W, Cin, cout, k, precision)
            else:
                # tf32: correlation of gy with the flipped kernel through the transposed packing, contraction over Cout
                wt, wt_lo = _pack_weight(_pad_channels(weight.detach().contiguous(), Cin), True, precision)
                gx = _conv_raw(g_hi, g_lo, wt, wt_lo, None, Cin, k, precision)
        if ctx.needs_input_grad[1]:
            # weight gradient: reduction over pixels, operands read in place as MN-major tiles
            d = ConvDesc(B, H, W, Cin, cout, k, precision, 0, 0)
            if ctx.nat is not None:
                # [Cout][tap][kpad] IS the stored layout: accumulate into the flat gradient buffer, nothing to return
                with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                    _lib.check(lib.pn_conv2d_wgrad(ctypes.byref(d), _lib.ptr(x_hi), _p(x_lo), _lib.ptr(g_hi), _p(g_lo),
                                                   _lib.ptr(ctx.nat.grad_flat), _lib.ptr(error_flag()), _stream()), "pn_conv2d_wgrad")
                    ctx.nat.owner.grad_ready(ctx.nat)      # data parallel: may start the all-reduce of a completed bucket (behind this launch)
                if side is not None:
                    for t in (x_hi, x_lo, g_hi, g_lo):     # read by the side stream after this node returned them to the allocator
                        if t is not None:
                            t.record_stream(side)
            else:
                # wgrad_side (the caller vouches that the ONLY consumer of this gradient waits for the side stream itself --
                # folded._FoldSetCUDA.backward does): same side stream, tensors allocated there
                with (torch.cuda.stream(side) if side is not None else contextlib.nullcontext()):
                    n = ctypes.c_size_t(0)
                    _lib.check(lib.pn_conv2d_wgrad_packed_elems(cout, Cin, k, precision, ctypes.byref(n)), "wgrad_packed_elems")
                    dwp = torch.empty(int(n.value), dtype=torch.float32, device=gy.device)
                    _lib.check(lib.pn_conv2d_wgrad(ctypes.byref(d), _lib.ptr(x_hi), _p(x_lo), _lib.ptr(g_hi), _p(g_lo),
                                                   _lib.ptr(dwp), _lib.ptr(error_flag()), _stream()), "pn_conv2d_wgrad")
                    gw_full = torch.empty(cout, Cin, k, k, dtype=torch.float32, device=gy.device)
                    if _state["unpack_tiled"]:
                        _lib.check(lib.pn_conv2d_unpack_weight_grad_tiled(_lib.ptr(dwp), _lib.ptr(gw_full), cout, Cin, k,
                                                                          int(n.value) // (cout * k * k), _stream()),
                                   "pn_conv2d_unpack_weight_grad_tiled")
                    else:
                        _lib.check(lib.pn_conv2d_unpack_weight_grad(_lib.ptr(dwp), _lib.ptr(gw_full), cout, Cin, k, precision, _stream()),
                                   "pn_conv2d_unpack_weight_grad")
                    gw = gw_full[:, :cin_w].contiguous() if cin_w != Cin else gw_full
                if side is not None:
                    for t in (x_hi, x_lo, g_hi, g_lo):
                        if t is not None:
                            t.record_stream(side)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = _lookup_channel_sum(gy)
            if gb is None:
                gb = torch.empty(cout, dtype=torch.float32, device=gy.device)
                _lib.check(lib.pn_channel_sum(_lib.ptr(gy), _lib.ptr(gb), B * H * W, cout, _stream()), "pn_channel_sum")
        return gx, gw, gb, None


def conv2d(x, weight, bias=None, wgrad_side=False):
    """x: [B,H,W,C] with C a multiple of channel_align(); weight may have fewer input channels (zero-padded).
    wgrad_side: the weight gradient may be produced on the weight-gradient side stream (the caller's consumer waits for it)."""
    return _Conv2d.apply(x, weight, bias, wgrad_side)


def conv2d_im2col(x, weight, bias=None, conv=None, align=None):
    """The same convolution (stride 1, zero pad k//2, nn.Conv2d + ConstantPad2d of layers01.py:28-30,36) for a layer with
    FEW input channels, evaluated as ONE 1x1 convolution over the im2col tensor [B,H,W,align(Cin*k*k)].

    The tensor-core engine walks (tap, 64-channel chunk) items; with Cin = 3 every one of the 25 taps of PackNet01's first
    layer is a full-width MMA over 3 useful reduction lanes (measured 0.45 ms forward + 0.35 ms weight gradient at 28 and
    36 TFLOP/s, profiles/r01_layer_table.txt).  The im2col tensor has 75 -> 80 channels = two chunks of one 1x1 tap:
    12.5x fewer MMA cycles for one extra 157 MB copy.  x: [B,H,W,C>=Cin] NHWC (only the first Cin channels are read);
    `conv` is the convolution callable (default: this module's tensor-core conv2d; the CPU tests pass a PyTorch one)."""
    import torch.nn.functional as F
    conv = conv2d if conv is None else conv
    align = channel_align() if align is None else align
    B, H, W, _ = x.shape
    cout, cin, k, _ = weight.shape
    m = k // 2
    xp = F.pad(x[..., :cin], (0, 0, m, m, m, m))              # zero frame of k//2 pixels around the map
    patches = xp.unfold(1, k, 1).unfold(2, k, 1)                 # [B,H,W,cin,k,k] view: (c, ky, kx) as in weight[co]
    K = cin * k * k
    Kp = (K + align - 1) // align * align
    col = torch.empty(B, H, W, Kp, dtype=x.dtype, device=x.device)
    col[..., :K].view(B, H, W, cin, k, k).copy_(patches)
    if Kp > K:
        col[..., K:].zero_()
    w2 = F.pad(weight.reshape(cout, K), (0, Kp - K)).view(cout, Kp, 1, 1)
    return conv(col, w2, bias)


# The GroupNorm backward already reduces its dx over the pixels; the convolution that consumes dx as its output
# gradient takes its bias gradient from here instead of re-reading the tensor.  An entry is valid only while the
# producing tensor object is alive (weak reference: no other tensor can own that storage) and has not been
# written since (autograd version counter).
_channel_sums = []


def _remember_channel_sum(dx, dsum):
    _channel_sums.append((weakref.ref(dx), dx.data_ptr(), tuple(dx.shape), dx._version, dsum))
    del _channel_sums[:-8]


def _lookup_channel_sum(g):
    for ref, ptr, shape, version, dsum in reversed(_channel_sums):
        # the SAME tensor object, unmodified since the GroupNorm backward wrote it: a second consumer's gradient accumulated
        # in place by autograd (or a hook editing dx) bumps the version counter and the cached sum is not used
        if ref() is not None and ptr == g.data_ptr() and shape == tuple(g.shape) and g._version == version and g.is_contiguous():
            return dsum
    return None


class _GroupNormELU(torch.autograd.Function):
    """y = ELU(GroupNorm16(x [+ x2])) (layers01.py:31-32,37 and :61-62,72)."""

    @staticmethod
    def forward(ctx, x, x2, gamma, beta, eps):
        _lib.require_f32(x, gamma, beta)
        x = x.contiguous()
        x2c = x2.contiguous() if x2 is not None else None
        B, H, W, C = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(B * 16 * 3, dtype=torch.float64, device=x.device)   # sums (double) + mean/rstd (float)
        ctx.emit = _state["precision"] == PRECISION_BF16X3 and _state["gn_emit_split"]
        if ctx.emit:
            # the consumer of y is (almost always) a tensor-core convolution: write its bf16 operand pair in the same pass
            hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
            lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
            _lib.check(_lib.lib().pn_groupnorm_elu_forward_split(_lib.ptr(x), _p(x2c), _lib.ptr(gamma.detach().contiguous()),
                                                                 _lib.ptr(beta.detach().contiguous()), float(eps), _lib.ptr(y),
                                                                 _lib.ptr(hi), _lib.ptr(lo), _lib.ptr(stats), B, H * W, C, _stream()),
                       "pn_groupnorm_elu_forward_split")
            _attach_split(y, hi, lo)
        else:
            _lib.check(_lib.lib().pn_groupnorm_elu_forward(_lib.ptr(x), _p(x2c), _lib.ptr(gamma.detach().contiguous()),
                                                           _lib.ptr(beta.detach().contiguous()), float(eps), _lib.ptr(y), None,
                                                           _lib.ptr(stats), B, H * W, C, C, 0, _stream()),
                       "pn_groupnorm_elu_forward")
        ctx.save_for_backward(x, x2c, y, gamma, stats)
        ctx.eps = float(eps)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, x2, y, gamma, stats = ctx.saved_tensors
        gy = gy.contiguous()
        B, H, W, C = x.shape
        dx = torch.empty_like(x)
        # dgamma | dbeta | dsum back to back: the cluster kernel zeroes the three with ONE memset
        small = torch.empty(3, C, dtype=torch.float32, device=x.device)
        dgamma, dbeta, dsum = small[0], small[1], small[2]
        bc = torch.empty(2 * C * B + 16 * B, dtype=torch.float64, device=x.device)   # doubles + float scratch tail
        if ctx.emit:
            # dx is the output gradient of the convolution that produced x: its bf16 pair comes out of the same pass
            hi = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
            lo = torch.empty(x.shape, dtype=torch.bfloat16, device=x.device)
            _lib.check(_lib.lib().pn_groupnorm_elu_backward_split(_lib.ptr(x), _p(x2), _lib.ptr(y), _lib.ptr(gy),
                                                                  _lib.ptr(gamma.detach().contiguous()), ctx.eps, _lib.ptr(stats),
                                                                  _lib.ptr(bc), _lib.ptr(dx), _lib.ptr(hi), _lib.ptr(lo),
                                                                  _lib.ptr(dgamma), _lib.ptr(dbeta), _lib.ptr(dsum), B, H * W, C,
                                                                  _stream()), "pn_groupnorm_elu_backward_split")
            _attach_split(dx, hi, lo)
        else:
            _lib.check(_lib.lib().pn_groupnorm_elu_backward(_lib.ptr(x), _p(x2), _lib.ptr(y), _lib.ptr(gy),
                                                            _lib.ptr(gamma.detach().contiguous()), ctx.eps, _lib.ptr(stats),
                                                            _lib.ptr(bc), _lib.ptr(dx), None, _lib.ptr(dgamma), _lib.ptr(dbeta),
                                                            _lib.ptr(dsum), B, H * W, C, C, 0, C, 0, _stream()),
                       "pn_groupnorm_elu_backward")
        _remember_channel_sum(dx, dsum)
        return dx, (dx if x2 is not None else None), dgamma, dbeta, None


def groupnorm_elu(x, gamma, beta, eps=1e-5, x2=None):
    return _GroupNormELU.apply(x, x2, gamma, beta, eps)


class _FeatureStencil(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w3, b3, pack, wgrad_side=False):
        _lib.require_f32(x, w3, b3)
        x = x.contiguous()
        B = x.shape[0]
        if pack:
            h, w, C = x.shape[1] // 2, x.shape[2] // 2, x.shape[3]
            out = torch.empty(B, h, w, 32 * C, dtype=torch.float32, device=x.device)
            cs = 32 * C
        else:
            h, w, C = x.shape[1], x.shape[2], x.shape[3]
            out = torch.empty(B, 2 * h, 2 * w, 2 * C, dtype=torch.float32, device=x.device)
            cs = 2 * C
        w3c = w3.detach().contiguous()
        _lib.check(_lib.lib().pn_feature_stencil_forward(int(pack), _lib.ptr(x), _lib.ptr(w3c), _lib.ptr(b3.detach().contiguous()),
                                                         _lib.ptr(out), None, B, h, w, C, cs, 0, _stream()),
                   "pn_feature_stencil_forward")
        ctx.save_for_backward(x, w3c)
        ctx.pack, ctx.dims, ctx.cs = bool(pack), (B, h, w, C), cs
        ctx.wgrad_side = bool(wgrad_side)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w3 = ctx.saved_tensors
        B, h, w, C = ctx.dims
        # The gradient usually arrives as a CHANNEL WINDOW of the gradient of a concatenation (torch.cat's backward hands out
        # narrow() views): the kernels read it in place through their channel stride -- no contiguous copy of the window
        cs = ctx.cs
        if (not g.is_contiguous() and g.dim() == 4 and g.stride(3) == 1 and g.stride(2) % 4 == 0 and g.stride(2) >= g.shape[3]
                and g.stride(1) == g.shape[2] * g.stride(2) and g.stride(0) == g.shape[1] * g.stride(1) and g.data_ptr() % 16 == 0):
            cs = g.stride(2)
        else:
            g = g.contiguous()
        gin = torch.empty_like(x)
        gw3 = torch.empty(216, dtype=torch.float32, device=x.device)
        gb3 = torch.empty(8, dtype=torch.float32, device=x.device)
        lib = _lib.lib()
        if ctx.wgrad_side and _state["wgrad_stream"] and g.is_cuda:
            # The Conv3d's weight / bias gradient on the weight-gradient side stream.  The caller vouched (at forward time) that
            # both parameters have no gradient yet: autograd's AccumulateGrad then only STORES these tensors -- no kernel reads
            # them before the end-of-backward join (_wgrad_side_stream).
            cur = torch.cuda.current_stream()
            side = _wgrad_side_stream(cur)
            ready = torch.cuda.Event()
            ready.record(cur)
            side.wait_event(ready)
            _lib.check(lib.pn_feature_stencil_backward_parts(int(ctx.pack), _lib.ptr(x), _lib.ptr(g), _lib.ptr(w3), _lib.ptr(gin),
                                                             None, None, B, h, w, C, cs, 0, 1, _stream()), "pn_feature_stencil_backward_parts")
            with torch.cuda.stream(side):
                _lib.check(lib.pn_feature_stencil_backward_parts(int(ctx.pack), _lib.ptr(x), _lib.ptr(g), _lib.ptr(w3), None,
                                                                 _lib.ptr(gw3), _lib.ptr(gb3), B, h, w, C, cs, 0, 2, _stream()),
                           "pn_feature_stencil_backward_parts")
            for t in (x, g, gw3, gb3):
                t.record_stream(side)
        else:
            _lib.check(lib.pn_feature_stencil_backward(int(ctx.pack), _lib.ptr(x), _lib.ptr(g), _lib.ptr(w3), _lib.ptr(gin),
                                                       _lib.ptr(gw3), _lib.ptr(gb3), B, h, w, C, cs, 0, _stream()),
                       "pn_feature_stencil_backward")
        return gin, gw3.view(8, 1, 3, 3, 3), gb3, None, None


class _HeadConv(torch.autograd.Function):
    """Conv2d(C -> 1, 3x3, pad 1) on an NHWC map (InvDepth.conv1, layers01.py:110-116)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        _lib.require_f32(x, weight, bias)
        x = x.contiguous()
        B, H, W, C = x.shape
        wt = weight.detach().reshape(C, 9).t().contiguous()          # [9][C] tap-major
        y = torch.empty(B, H, W, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().pn_head_conv_forward(_lib.ptr(x), _lib.ptr(wt), _lib.ptr(bias.detach().contiguous()), _lib.ptr(y),
                                                   B, H, W, C, _stream()), "pn_head_conv_forward")
        ctx.save_for_backward(x, wt)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, wt = ctx.saved_tensors
        B, H, W, C = x.shape
        dy = dy.contiguous()
        dx = torch.empty_like(x)
        dw = torch.empty(9, C, dtype=torch.float32, device=x.device)
        db = torch.empty(1, dtype=torch.float32, device=x.device)
        _lib.check(_lib.lib().pn_head_conv_backward(_lib.ptr(x), _lib.ptr(dy), _lib.ptr(wt), _lib.ptr(dx), _lib.ptr(dw), _lib.ptr(db),
                                                    B, H, W, C, _stream()), "pn_head_conv_backward")
        return dx, dw.t().reshape(1, C, 3, 3), db


def head_conv(x, weight, bias):
    """[B,H,W,C] NHWC, weight [1,C,3,3], bias [1] -> [B,H,W]"""
    return _HeadConv.apply(x, weight, bias)


def _grads_unset(*params):
    """the parameters are leaves without a gradient so far (optimizer.zero_grad(set_to_none=True) ran): a gradient produced for
    them is only stored by autograd, never added to -- the condition for producing it on the weight-gradient side stream"""
    return all(p.is_leaf and p.grad is None for p in params)


def pack_features(x, w3, b3):
    """[B,2h,2w,C] -> [B,h,w,32C]"""
    return _FeatureStencil.apply(x, w3, b3, True, _grads_unset(w3, b3))


def unpack_features(u, w3, b3):
    """[B,h,w,Cu] -> [B,2h,2w,2Cu]"""
    return _FeatureStencil.apply(u, w3, b3, False, _grads_unset(w3, b3))
