"""PackNet01 depth network on the sm_100a kernels -- drop-in for
packnet_sfm/networks/depth/PackNet01.py::PackNet01 (same constructor, same forward contract, same 216
state-dict keys and shapes, same initialisation law) and the layer classes of
packnet_sfm/networks/layers/packnet/layers01.py.

The nn.Conv2d / nn.Conv3d / nn.GroupNorm members only HOLD the parameters (so checkpoints of the reference
load by name, packnet_sfm/utils/load.py:146-157); their own forward is never called.  Feature maps are NHWC
tensors; every convolution runs on the tcgen05 implicit-GEMM kernel, the Conv3d feature stencils and
GroupNorm+ELU on the fused HBM-bound kernels (packnet_sfm_b200/functional.py)."""
import functools

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import folded
from . import functional as PF


def _cat_channels(tensors):
    """torch.cat along channels (PackNet01.py:138-175) on NHWC maps, zero-padded to the channel multiple the
    following convolution's operand needs (16-byte TMA row pitch: 4 fp32 / 8 bf16 elements)."""
    c = sum(t.shape[-1] for t in tensors)
    pad = (-c) % PF.channel_align()
    if pad:
        b, h, w, _ = tensors[0].shape
        tensors = list(tensors) + [torch.zeros(b, h, w, pad, dtype=tensors[0].dtype, device=tensors[0].device)]
    return torch.cat(tensors, dim=-1)


class Conv2D(nn.Module):
    """2D convolution with GroupNorm and ELU (layers01.py:10-37)."""

    def __init__(self, in_channels, out_channels, kernel_size, stride):
        super().__init__()
        assert stride == 1
        self.kernel_size = kernel_size
        self.conv_base = nn.Conv2d(in_channels, out_channels, kernel_size=kernel_size, stride=stride)
        self.normalize = torch.nn.GroupNorm(16, out_channels)

    def forward(self, x):
        z = PF.conv2d(x, self.conv_base.weight, self.conv_base.bias)
        return PF.groupnorm_elu(z, self.normalize.weight, self.normalize.bias, self.normalize.eps)


class ResidualConv(nn.Module):
    """2D convolutional residual block with GroupNorm and ELU (layers01.py:40-72)."""

    def __init__(self, in_channels, out_channels, stride, dropout=None):
        super().__init__()
        assert stride == 1
        self.conv1 = Conv2D(in_channels, out_channels, 3, stride)
        self.conv2 = Conv2D(out_channels, out_channels, 3, 1)
        self.conv3 = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=stride)
        self.normalize = torch.nn.GroupNorm(16, out_channels)
        self.dropout = dropout
        if dropout:
            # same parameter names as the reference's nn.Sequential(conv3, Dropout2d) (layers01.py:64-65)
            self.conv3 = nn.Sequential(self.conv3, nn.Dropout2d(dropout))

    def forward(self, x):
        x_out = self.conv2(self.conv1(x))
        if self.dropout:
            conv3 = self.conv3[0]
            shortcut = PF.conv2d(x, conv3.weight, conv3.bias)
            # Dropout2d drops whole channels: broadcast a [B,1,1,C] mask over the NHWC map
            shortcut = self.conv3[1](shortcut.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
        else:
            shortcut = PF.conv2d(x, self.conv3.weight, self.conv3.bias)
        return PF.groupnorm_elu(x_out, self.normalize.weight, self.normalize.bias, self.normalize.eps, x2=shortcut)


def ResidualBlock(in_channels, out_channels, num_blocks, stride, dropout=None):
    """layers01.py:75-95."""
    layers = [ResidualConv(in_channels, out_channels, stride, dropout=dropout)]
    for _ in range(1, num_blocks):
        layers.append(ResidualConv(out_channels, out_channels, 1, dropout=dropout))
    return nn.Sequential(*layers)


class InvDepth(nn.Module):
    """Inverse depth head (layers01.py:98-122): pad 1 -> Conv2d(C->1, 3x3) -> sigmoid / min_depth.
    One output channel is HBM-bound fp32 work, not a GEMM: pn_head_conv_* (exact fp32 FMAs)."""

    def __init__(self, in_channels, out_channels=1, min_depth=0.5):
        super().__init__()
        self.min_depth = min_depth
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1)

    def forward(self, x):
        if x.shape[3] % 4 or self.conv1.out_channels != 1:
            raise NotImplementedError("InvDepth: the head kernel takes C %% 4 == 0 input channels and one output channel "
                                      "(got %d -> %d); there is no library fallback" % (x.shape[3], self.conv1.out_channels))
        y = PF.head_conv(x, self.conv1.weight, self.conv1.bias).unsqueeze(1)
        return torch.sigmoid(y) / self.min_depth          # [B,1,H,W] (NCHW == NHWC for one channel)


class PackLayerConv3d(nn.Module):
    """Packing layer with 3d convolutions (layers01.py:213-247)."""

    def __init__(self, in_channels, kernel_size, r=2, d=8):
        super().__init__()
        assert r == 2 and d == 8
        self.conv = Conv2D(in_channels * (r ** 2) * d, in_channels, kernel_size, 1)
        self.conv3d = nn.Conv3d(1, d, kernel_size=(3, 3, 3), stride=(1, 1, 1), padding=(1, 1, 1))

    def folds(self, h, w):
        """whether the layer is evaluated as one folded convolution on a packed map of h x w pixels (policy + geometry)"""
        k = self.conv.kernel_size
        return PF.pack_fold_enabled(h * w) and k in (3, 5) and min(h, w) >= 2 * (k // 2) + 1

    def prefold(self, w):
        """weight-only half of the folded evaluation for a packed map of width w, on the CURRENT stream (PackNet01.forward calls
        this under a side stream before the first layer runs)"""
        self._pre = (folded.prefold(self.conv.conv_base.weight, self.conv3d.weight, self.conv3d.bias, w), torch.cuda.Event())
        self._pre[1].record()

    def forward(self, x):
        h, w = x.shape[1] // 2, x.shape[2] // 2
        if self.folds(h, w):
            # conv3d and conv2d composed into one (k+2)x(k+2) convolution of the space-to-depth tensor + exact frame terms
            pre = None
            if getattr(self, "_pre", None) is not None:
                (folds, beta, dB), ev = self._pre
                self._pre = None
                cur = torch.cuda.current_stream()
                cur.wait_event(ev)
                for t in tuple(folds) + (beta, dB):      # allocated on the side stream, consumed here
                    t.record_stream(cur)
                pre = (folds, beta, dB)
            # with the folds on their own stream the effective weight's gradient has ONE consumer, the fold backward, which waits
            # for the weight-gradient side stream itself: the largest weight-gradient launch of the step leaves the critical path
            conv = functools.partial(PF.conv2d, wgrad_side=True) if pre is not None else PF.conv2d
            z = folded.pack_conv_folded(x, self.conv.conv_base.weight, self.conv.conv_base.bias, self.conv3d.weight,
                                        self.conv3d.bias, conv, pre=pre)
            return PF.groupnorm_elu(z, self.conv.normalize.weight, self.conv.normalize.bias, self.conv.normalize.eps)
        feats = PF.pack_features(x, self.conv3d.weight, self.conv3d.bias)
        return self.conv(feats)


class UnpackLayerConv3d(nn.Module):
    """Unpacking layer with 3d convolutions (layers01.py:250-286)."""

    def __init__(self, in_channels, out_channels, kernel_size, r=2, d=8):
        super().__init__()
        assert r == 2 and d == 8
        self.conv = Conv2D(in_channels, out_channels * (r ** 2) // d, kernel_size, 1)
        self.conv3d = nn.Conv3d(1, d, kernel_size=(3, 3, 3), stride=(1, 1, 1), padding=(1, 1, 1))

    def forward(self, x):
        u = self.conv(x)
        return PF.unpack_features(u, self.conv3d.weight, self.conv3d.bias)


def native_conv_weights(net, image_hw=None):
    """The convolution weights of `net` that reach the tensor-core engine AS PARAMETERS (functional.conv2d on the parameter
    itself): optim.FlatAdam stores these in the engine's [Cout][tap][kpad] layout.  Not in the list: InvDepth heads (SIMT
    kernels), the Conv3d stencils, the first layer when it runs through its im2col re-composition, and the Conv2d of a pack
    layer that is evaluated FOLDED at the given image size (its weight is an operand of the fold kernels, in OIHW)."""
    out = []
    folded_convs = set()
    if isinstance(net, PackNet01) and image_hw is not None:
        H, W = image_hw
        for i, name in enumerate(("pack1", "pack2", "pack3", "pack4", "pack5")):
            layer = getattr(net, name)
            h, w = H >> (i + 1), W >> (i + 1)
            if layer.folds(h, w):
                folded_convs.add(id(layer.conv))
    skip_first = isinstance(net, PackNet01) and PF.im2col_first_enabled()
    for m in net.modules():
        if isinstance(m, Conv2D) and id(m) not in folded_convs and not (skip_first and m is net.pre_calc):
            out.append(m.conv_base.weight)
        elif isinstance(m, ResidualConv):
            out.append((m.conv3[0] if m.dropout else m.conv3).weight)
    return out


def gradient_buckets(net, image_hw=None):
    """Stored weights grouped by WHEN their gradients complete in the backward (reverse forward order, PackNet01.py:106-176):
    [decoder: iconv1 .. unpack5] then [pack5, conv5, pack4: 110 M of the 128 M parameters, complete while the high-resolution
    half of the encoder's backward is still to run].  optim.FlatAdam launches each bucket's all-reduce from the backward."""
    native = {id(p) for p in native_conv_weights(net, image_hw)}
    if not isinstance(net, PackNet01):
        return []

    def weights(*mods):
        out = []
        for m in mods:
            for sub in m.modules():
                for p in sub.parameters(recurse=False):
                    if id(p) in native and all(p is not q for q in out):
                        out.append(p)
        return out

    decoder = weights(net.iconv1, net.unpack1, net.iconv2, net.unpack2, net.iconv3, net.unpack3, net.iconv4, net.unpack4,
                      net.iconv5, net.unpack5)
    deep = weights(net.pack5, net.conv5, net.pack4)
    return [decoder, deep]


class PackNet01(nn.Module):
    """PackNet network with 3d convolutions (version 01, from the CVPR paper) -- PackNet01.py:7-185.

    forward(rgb [B,3,H,W]) -> {'inv_depths': [disp1..disp4]} in train mode (H, H/2, H/4, H/8; NCHW
    [B,1,h,w]) and {'inv_depths': disp1} in eval mode, as the reference does (PackNet01.py:178-185)."""

    def __init__(self, dropout=None, version=None, **kwargs):
        super().__init__()
        self.version = version[1:]
        in_channels, out_channels = 3, 1
        ni, no = 64, out_channels
        n1, n2, n3, n4, n5 = 64, 64, 128, 256, 512
        num_blocks = [2, 2, 3, 3]
        pack_kernel = [5, 3, 3, 3, 3]
        unpack_kernel = [3, 3, 3, 3, 3]
        iconv_kernel = [3, 3, 3, 3, 3]
        self.pre_calc = Conv2D(in_channels, ni, 5, 1)
        if self.version == 'A':
            n1o, n1i = n1, n1 + ni + no
            n2o, n2i = n2, n2 + n1 + no
            n3o, n3i = n3, n3 + n2 + no
            n4o, n4i = n4, n4 + n3
            n5o, n5i = n5, n5 + n4
        elif self.version == 'B':
            n1o, n1i = n1, n1 + no
            n2o, n2i = n2, n2 + no
            n3o, n3i = n3 // 2, n3 // 2 + no
            n4o, n4i = n4 // 2, n4 // 2
            n5o, n5i = n5 // 2, n5 // 2
        else:
            raise ValueError('Unknown PackNet version {}'.format(version))

        self.pack1 = PackLayerConv3d(n1, pack_kernel[0])
        self.pack2 = PackLayerConv3d(n2, pack_kernel[1])
        self.pack3 = PackLayerConv3d(n3, pack_kernel[2])
        self.pack4 = PackLayerConv3d(n4, pack_kernel[3])
        self.pack5 = PackLayerConv3d(n5, pack_kernel[4])

        self.conv1 = Conv2D(ni, n1, 7, 1)
        self.conv2 = ResidualBlock(n1, n2, num_blocks[0], 1, dropout=dropout)
        self.conv3 = ResidualBlock(n2, n3, num_blocks[1], 1, dropout=dropout)
        self.conv4 = ResidualBlock(n3, n4, num_blocks[2], 1, dropout=dropout)
        self.conv5 = ResidualBlock(n4, n5, num_blocks[3], 1, dropout=dropout)

        self.unpack5 = UnpackLayerConv3d(n5, n5o, unpack_kernel[0])
        self.unpack4 = UnpackLayerConv3d(n5, n4o, unpack_kernel[1])
        self.unpack3 = UnpackLayerConv3d(n4, n3o, unpack_kernel[2])
        self.unpack2 = UnpackLayerConv3d(n3, n2o, unpack_kernel[3])
        self.unpack1 = UnpackLayerConv3d(n2, n1o, unpack_kernel[4])

        self.iconv5 = Conv2D(n5i, n5, iconv_kernel[0], 1)
        self.iconv4 = Conv2D(n4i, n4, iconv_kernel[1], 1)
        self.iconv3 = Conv2D(n3i, n3, iconv_kernel[2], 1)
        self.iconv2 = Conv2D(n2i, n2, iconv_kernel[3], 1)
        self.iconv1 = Conv2D(n1i, n1, iconv_kernel[4], 1)

        # parameter-free members kept for attribute parity with the reference (PackNet01.py:86-89)
        self.unpack_disps = nn.PixelShuffle(2)
        self.unpack_disp4 = nn.Upsample(scale_factor=2, mode='nearest', align_corners=None)
        self.unpack_disp3 = nn.Upsample(scale_factor=2, mode='nearest', align_corners=None)
        self.unpack_disp2 = nn.Upsample(scale_factor=2, mode='nearest', align_corners=None)

        self.disp4_layer = InvDepth(n4, out_channels=out_channels)
        self.disp3_layer = InvDepth(n3, out_channels=out_channels)
        self.disp2_layer = InvDepth(n2, out_channels=out_channels)
        self.disp1_layer = InvDepth(n1, out_channels=out_channels)

        self.init_weights()

    def init_weights(self):
        """Xavier-uniform conv weights, zero biases (PackNet01.py:98-104)."""
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.Conv3d)):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    m.bias.data.zero_()

    @staticmethod
    def _up_nhwc(disp):
        """nearest x2 of a [B,1,h,w] map, returned as NHWC [B,2h,2w,1] (nn.Upsample, PackNet01.py:87-89)."""
        up = F.interpolate(disp, scale_factor=2, mode='nearest')
        return up.permute(0, 2, 3, 1)

    def forward(self, rgb):
        _lib.require_f32(rgb)          # CUDA fp32 only: there is no CPU path and no silent dtype reinterpretation
        B, _, H, W = rgb.shape
        if H % 32 or W % 32:
            raise ValueError("PackNet01 needs H and W divisible by 32 (got %dx%d)" % (H, W))
        # NCHW image -> NHWC, zero-padded to the operand's channel multiple (TMA row pitch)
        cpad = PF.channel_align() - 3
        x_in = torch.cat([rgb.permute(0, 2, 3, 1), torch.zeros(B, H, W, cpad, dtype=rgb.dtype, device=rgb.device)], -1)
        if PF.prefold_stream_enabled() and rgb.is_cuda:
            # the weight folds of the folded pack layers depend on parameters only: a side stream computes them while the first
            # layers run (autograd replays their backward on that stream too; a captured step keeps the fork / join)
            cur = torch.cuda.current_stream()
            if getattr(self, "_fold_stream", None) is None:
                self._fold_stream = torch.cuda.Stream()
            self._fold_stream.wait_stream(cur)
            with torch.cuda.stream(self._fold_stream):
                for i, name in enumerate(("pack1", "pack2", "pack3", "pack4", "pack5")):
                    layer = getattr(self, name)
                    if layer.folds(H >> (i + 1), W >> (i + 1)):
                        layer.prefold(W >> (i + 1))
        if PF.im2col_first_enabled():
            # staged: the 3 -> 64 5x5 layer as one 1x1 convolution over its im2col tensor (functional.conv2d_im2col)
            pc = self.pre_calc
            z = PF.conv2d_im2col(rgb.permute(0, 2, 3, 1), pc.conv_base.weight, pc.conv_base.bias)
            x = PF.groupnorm_elu(z, pc.normalize.weight, pc.normalize.bias, pc.normalize.eps)
        else:
            x = self.pre_calc(x_in.contiguous())

        x1 = self.conv1(x)
        x1p = self.pack1(x1)
        x2 = self.conv2(x1p)
        x2p = self.pack2(x2)
        x3 = self.conv3(x2p)
        x3p = self.pack3(x3)
        x4 = self.conv4(x3p)
        x4p = self.pack4(x4)
        x5 = self.conv5(x4p)
        x5p = self.pack5(x5)

        skip1, skip2, skip3, skip4, skip5 = x, x1p, x2p, x3p, x4p
        A = self.version == 'A'

        unpack5 = self.unpack5(x5p)
        concat5 = _cat_channels((unpack5, skip5)) if A else unpack5 + skip5
        iconv5 = self.iconv5(concat5)

        unpack4 = self.unpack4(iconv5)
        concat4 = _cat_channels((unpack4, skip4)) if A else unpack4 + skip4
        iconv4 = self.iconv4(concat4)
        disp4 = self.disp4_layer(iconv4)
        udisp4 = self._up_nhwc(disp4)

        unpack3 = self.unpack3(iconv4)
        concat3 = _cat_channels((unpack3, skip3, udisp4)) if A else _cat_channels((unpack3 + skip3, udisp4))
        iconv3 = self.iconv3(concat3)
        disp3 = self.disp3_layer(iconv3)
        udisp3 = self._up_nhwc(disp3)

        unpack2 = self.unpack2(iconv3)
        concat2 = _cat_channels((unpack2, skip2, udisp3)) if A else _cat_channels((unpack2 + skip2, udisp3))
        iconv2 = self.iconv2(concat2)
        disp2 = self.disp2_layer(iconv2)
        udisp2 = self._up_nhwc(disp2)

        unpack1 = self.unpack1(iconv2)
        concat1 = _cat_channels((unpack1, skip1, udisp2)) if A else _cat_channels((unpack1 + skip1, udisp2))
        iconv1 = self.iconv1(concat1)
        disp1 = self.disp1_layer(iconv1)

        if self.training:
            return {'inv_depths': [disp1, disp2, disp3, disp4]}
        return {'inv_depths': disp1}
