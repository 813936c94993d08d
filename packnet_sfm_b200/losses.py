"""MultiViewPhotometricLoss on the fused sm_100a kernels -- drop-in for
packnet_sfm/losses/multiview_photometric_loss.py::MultiViewPhotometricLoss (same constructor keywords,
same forward signature, same return dict) and packnet_sfm/losses/loss_base.py::{LossBase,ProgressiveScaling}.

The whole forward (warp, SSIM, L1, auto-mask min, smoothness, reduction) is ONE tile kernel plus a small
prep kernel; the backward is one more launch of the same tile program (include/packnet_b200.h:
pn_loss_forward / pn_loss_backward).  Gradients flow to `inv_depths[i]` and `poses[j].mat`."""
import ctypes
import os

import numpy as np
import torch
import torch.nn as nn

from . import _lib


class ProgressiveScaling:
    """loss_base.py:9-48 -- after `progressive_scaling * (i+1)` training progress drop one scale."""

    def __init__(self, progressive_scaling, num_scales=4):
        self.num_scales = num_scales
        if progressive_scaling > 0.0:
            self.progressive_scaling = np.float32(
                [progressive_scaling * (i + 1) for i in range(num_scales - 1)] + [1.0])
        else:
            self.progressive_scaling = progressive_scaling

    def __call__(self, progress):
        if isinstance(self.progressive_scaling, np.ndarray):
            return int(self.num_scales - np.searchsorted(self.progressive_scaling, progress))
        return self.num_scales


class LossBase(nn.Module):
    """loss_base.py:52-74."""

    def __init__(self):
        super().__init__()
        self._logs = {}
        self._metrics = {}

    @property
    def logs(self):
        return self._logs

    @property
    def metrics(self):
        return self._metrics

    def add_metric(self, key, val):
        self._metrics[key] = val.detach()


_grouped = os.environ.get("PN_LOSS_GROUPED", "1") == "1"     # default since round 2 (B200: fwd 0.193 -> 0.145, bwd 0.599 -> 0.436 ms)


_fused_training = os.environ.get("PN_LOSS_FUSED", "1") == "1"


def set_fused_training(on):
    """One tile launch for the loss AND its unit gradients when an input requires a gradient (pn_loss_forward_backward; grouped
    program only); off = forward launch + backward launch of the tile program (the round-1 call sequence).  Returns the
    previous setting."""
    global _fused_training
    prev, _fused_training = _fused_training, bool(on)
    return prev


def set_grouped_kernel(on):
    """Select the grouped-scale tile program (csrc/loss_group_kernel.cuh, PN_LOSS_FLAG_GROUPED) for the descriptors built
    from now on (the default since round 2; PN_LOSS_GROUPED=0 in the environment or set_grouped_kernel(False) selects the
    first-generation tile program, which stays tested).  Returns the previous setting."""
    global _grouped
    prev, _grouped = _grouped, bool(on)
    return prev


def _make_desc(B, H, W, num_context, shapes, ssim_w, smooth_w, C1, C2, reduce_min, automask, nearest_upsample=False):
    d = _lib.LossDesc()
    d.flags = _lib.PN_LOSS_FLAG_GROUPED if _grouped else 0
    d.batch, d.height, d.width = B, H, W
    d.num_context, d.num_scales = num_context, len(shapes)
    for i, (h, w) in enumerate(shapes):
        if nearest_upsample and (h, w) != (H, W):
            # the map stays at its own resolution and is read as upsample_output(mode='nearest') would have made it
            sh = (H // h).bit_length() - 1
            if h << sh != H or w << sh != W:
                raise ValueError("nearest_upsample: %dx%d is not a power-of-two reduction of %dx%d" % (h, w, H, W))
            d.scale_h[i], d.scale_w[i], d.inv_shift[i] = H, W, sh
            continue
        d.scale_h[i], d.scale_w[i] = h, w
    d.ssim_loss_weight, d.smooth_loss_weight, d.C1, d.C2 = ssim_w, smooth_w, C1, C2
    d.reduce_min, d.automask = int(reduce_min), int(automask)
    return d


class _FusedLoss(torch.autograd.Function):
    """tensors = context[N] + inv_depths[n] + pose matrices[N]; returns out[4] = (loss, metrics..)."""

    @staticmethod
    def forward(ctx, cfg, image, K, ref_K, *tensors):
        N, n = cfg["N"], cfg["n"]
        context = [t.contiguous() for t in tensors[:N]]
        inv = [t.contiguous() for t in tensors[N:N + n]]
        poses = [t.contiguous() for t in tensors[N + n:]]
        image, K, ref_K = image.contiguous(), K.contiguous(), ref_K.contiguous()
        _lib.require_f32(image, K, ref_K, *context, *inv, *poses)
        B, _, H, W = image.shape
        desc = _make_desc(B, H, W, N, [tuple(d.shape[-2:]) for d in inv], cfg["ssim_w"], cfg["smooth_w"],
                          cfg["C1"], cfg["C2"], cfg["reduce_min"], cfg["automask"], cfg.get("nearest_upsample", False))
        lib = _lib.lib()
        nbytes = ctypes.c_size_t(0)
        _lib.check(lib.pn_loss_workspace_bytes(ctypes.byref(desc), ctypes.byref(nbytes)), "pn_loss_workspace_bytes")
        ws = torch.empty(max(int(nbytes.value), 16), dtype=torch.uint8, device=image.device)
        out = torch.empty(4, dtype=torch.float32, device=image.device)
        ctx.desc, ctx.ws, ctx.N, ctx.n = desc, ws, N, n
        ctx.saved = (image, K, ref_K, context, inv, poses)
        ctx.unit = None
        if _fused_training and (desc.flags & _lib.PN_LOSS_FLAG_GROUPED) and any(ctx.needs_input_grad[4 + N:]):
            # training: ONE tile launch gives the loss and the unit gradients (pn_loss_forward_backward); backward() scales them.
            # All unit gradients live in one buffer: the library zeroes it with a single memset.
            sizes = [d.numel() for d in inv] + [p.numel() for p in poses]
            offs = np.cumsum([0] + [(s + 3) // 4 * 4 for s in sizes])
            flat = torch.empty(int(offs[-1]), dtype=torch.float32, device=image.device)
            unit = [flat[int(o):int(o) + s].view(t.shape) for o, s, t in zip(offs[:-1], sizes, inv + poses)]
            _lib.check(lib.pn_loss_forward_backward(ctypes.byref(desc), _lib.ptr(image), _lib.ptr_array(context),
                                                    _lib.ptr_array(inv), _lib.ptr(K), _lib.ptr(ref_K), _lib.ptr_array(poses),
                                                    _lib.ptr(out), _lib.ptr_array(unit[:n]), _lib.ptr_array(unit[n:]),
                                                    flat.numel() * 4, _lib.ptr(ws), ws.numel(), _lib.current_stream()),
                       "pn_loss_forward_backward")
            ctx.unit = (flat, unit)
            return out
        _lib.check(lib.pn_loss_forward(ctypes.byref(desc), _lib.ptr(image), _lib.ptr_array(context),
                                       _lib.ptr_array(inv), _lib.ptr(K), _lib.ptr(ref_K), _lib.ptr_array(poses),
                                       _lib.ptr(out), _lib.ptr(ws), ws.numel(), _lib.current_stream()),
                   "pn_loss_forward")
        return out

    @staticmethod
    def backward(ctx, grad_out):
        image, K, ref_K, context, inv, poses = ctx.saved
        lib = _lib.lib()
        g = grad_out.contiguous().to(torch.float32)
        ginv = [torch.empty_like(d) for d in inv]
        gpose = [torch.empty_like(p) for p in poses]
        if ctx.unit is not None:
            unit = ctx.unit[1]
            _lib.check(lib.pn_loss_backward_finish(ctypes.byref(ctx.desc), _lib.ptr(g), _lib.ptr_array(unit[:ctx.n]),
                                                   _lib.ptr_array(unit[ctx.n:]), _lib.ptr_array(ginv), _lib.ptr_array(gpose),
                                                   _lib.ptr(ctx.ws), ctx.ws.numel(), _lib.current_stream()),
                       "pn_loss_backward_finish")
            return (None, None, None, None) + (None,) * ctx.N + tuple(ginv) + tuple(gpose)
        _lib.check(lib.pn_loss_backward(ctypes.byref(ctx.desc), _lib.ptr(image), _lib.ptr_array(context),
                                        _lib.ptr_array(inv), _lib.ptr(K), _lib.ptr(ref_K), _lib.ptr_array(poses),
                                        _lib.ptr(g), _lib.ptr_array(ginv), _lib.ptr_array(gpose), _lib.ptr(ctx.ws),
                                        ctx.ws.numel(), _lib.current_stream()),
                   "pn_loss_backward")
        return (None, None, None, None) + (None,) * ctx.N + tuple(ginv) + tuple(gpose)


class MultiViewPhotometricLoss(LossBase):
    """Self-supervised multiview photometric loss (reference: multiview_photometric_loss.py:57-344).

    Constructor keywords, their defaults and the assertion are the reference's (:91-114); the remaining
    keys of `config.model.loss` are swallowed by **kwargs exactly as there.  Differences, all loud:
      * clip_loss > 0 (a host-synchronising mean/std clamp, :218-221) -> NotImplementedError
      * padding_mode != 'zeros'                                        -> NotImplementedError
      * ssim_loss_weight <= 0 (3-channel L1 branch, :216-217)          -> RuntimeError from the library
    `occ_reg_weight` and `disp_norm` are accepted and unused, as in the reference."""

    def __init__(self, num_scales=4, ssim_loss_weight=0.85, occ_reg_weight=0.1, smooth_loss_weight=0.1,
                 C1=1e-4, C2=9e-4, photometric_reduce_op='mean', disp_norm=True, clip_loss=0.5,
                 progressive_scaling=0.0, padding_mode='zeros', automask_loss=False, **kwargs):
        super().__init__()
        self.n = num_scales
        self.ssim_loss_weight = ssim_loss_weight
        self.occ_reg_weight = occ_reg_weight
        self.smooth_loss_weight = smooth_loss_weight
        self.C1 = C1
        self.C2 = C2
        self.photometric_reduce_op = photometric_reduce_op
        self.disp_norm = disp_norm
        self.clip_loss = clip_loss
        self.padding_mode = padding_mode
        self.automask_loss = automask_loss
        self.progressive_scaling = ProgressiveScaling(progressive_scaling, self.n)
        if self.automask_loss:
            assert self.photometric_reduce_op == 'min', \
                'For automasking only the min photometric_reduce_op is supported.'

    @property
    def logs(self):
        return {'num_scales': self.n}

    def forward(self, image, context, inv_depths, K, ref_K, poses, return_logs=False, progress=0.0, nearest_upsample=False):
        """nearest_upsample=True (not in the reference's signature; its callers never pass it): `inv_depths` are the network's
        maps at H, H/2, H/4, H/8 and the loss reads them as SfmModel's upsample_output(mode='nearest') would have delivered
        them (models/model_utils.py:152-180, SfmModel.py:87-88) -- index >> s in the kernel's load, the 2^s x 2^s block sum in
        its backward -- instead of three full-resolution copies and their backward passes."""
        if self.clip_loss > 0.0:
            raise NotImplementedError("clip_loss > 0 is not implemented in the fused kernel (training default is 0.0, "
                                      "configs/default_config.py:99)")
        if self.padding_mode != 'zeros':
            raise NotImplementedError("padding_mode=%r: only 'zeros' is implemented" % (self.padding_mode,))
        if self.photometric_reduce_op not in ('min', 'mean'):
            raise NotImplementedError('Unknown photometric_reduce_op: {}'.format(self.photometric_reduce_op))
        self.n = self.progressive_scaling(progress)
        n = self.n
        mats = [p.mat if hasattr(p, "mat") else p for p in poses]
        cfg = dict(N=len(context), n=n, ssim_w=float(self.ssim_loss_weight), smooth_w=float(self.smooth_loss_weight),
                   C1=float(self.C1), C2=float(self.C2), reduce_min=self.photometric_reduce_op == 'min',
                   automask=bool(self.automask_loss), nearest_upsample=bool(nearest_upsample))
        out = _FusedLoss.apply(cfg, image, K.float(), ref_K.float(), *context, *inv_depths[:n], *mats)
        self.add_metric('photometric_loss', out[1])
        if self.smooth_loss_weight > 0.0:
            self.add_metric('smoothness_loss', out[2])
        return {'loss': out[0:1], 'metrics': self.metrics}


def warp_tap_indices(inv_depth, K, ref_K, pose, full_width=None):
    """Inspection hook (pn_loss_warp_indices): integer tap origins int32 [B,h,w,2] and float coordinates
    [B,h,w,2] for one scale / one context, from the device function the loss kernels use."""
    inv_depth, K, ref_K = inv_depth.contiguous(), K.contiguous().float(), ref_K.contiguous().float()
    pose = (pose.mat if hasattr(pose, "mat") else pose).contiguous()
    _lib.require_f32(inv_depth, K, ref_K, pose)
    B, _, h, w = inv_depth.shape
    W = w if full_width is None else full_width
    H = h if full_width is None else h * (W // w)
    desc = _make_desc(B, H, W, 1, [(h, w)], 0.85, 0.0, 1e-4, 9e-4, True, False)
    lib = _lib.lib()
    nbytes = ctypes.c_size_t(0)
    _lib.check(lib.pn_loss_workspace_bytes(ctypes.byref(desc), ctypes.byref(nbytes)), "pn_loss_workspace_bytes")
    ws = torch.empty(int(nbytes.value), dtype=torch.uint8, device=inv_depth.device)
    taps = torch.empty(B, h, w, 2, dtype=torch.int32, device=inv_depth.device)
    coords = torch.empty(B, h, w, 2, dtype=torch.float32, device=inv_depth.device)
    _lib.check(lib.pn_loss_warp_indices(ctypes.byref(desc), 0, _lib.ptr(inv_depth), _lib.ptr(K), _lib.ptr(ref_K),
                                        _lib.ptr(pose), _lib.ptr(taps), _lib.ptr(coords), _lib.ptr(ws), ws.numel(),
                                        _lib.current_stream()), "pn_loss_warp_indices")
    return taps, coords
