"""Host-side (PyTorch) glue around the two kernel-backed classes, mirroring the reference's model layer so
the self-supervised step can run where /root/reference is absent (the GPU box):

  PoseNet       <- packnet_sfm/networks/pose/PoseNet.py:38-86   (north_star: PoseNet stays PyTorch host code)
  SelfSupModel  <- packnet_sfm/models/SfmModel.py:53-127 + SelfSupModel.py:63-97 (flip, upsample, pose, loss)

With the reference on PYTHONPATH, packnet_sfm_b200.dropin.install() makes the reference's own SfmModel /
SelfSupModel / ModelWrapper use PackNet01 and MultiViewPhotometricLoss from this package unchanged."""
import os
import random

import torch
import torch.nn as nn
import torch.nn.functional as F

from .geometry import Pose
from .losses import MultiViewPhotometricLoss
from .networks import PackNet01


def conv_gn(in_planes, out_planes, kernel_size=3):
    """PoseNet.py:11-34: Conv2d(stride 2) + GroupNorm(16) + ReLU."""
    return nn.Sequential(
        nn.Conv2d(in_planes, out_planes, kernel_size=kernel_size, padding=(kernel_size - 1) // 2, stride=2),
        nn.GroupNorm(16, out_planes),
        nn.ReLU(inplace=True))


class PoseNet(nn.Module):
    """Pose network (PoseNet.py:38-86), same parameter names; 7 strided conv blocks + 1x1 head, x0.01."""

    def __init__(self, nb_ref_imgs=2, rotation_mode='euler', **kwargs):
        super().__init__()
        self.nb_ref_imgs = nb_ref_imgs
        self.rotation_mode = rotation_mode
        ch = [16, 32, 64, 128, 256, 256, 256]
        self.conv1 = conv_gn(3 * (1 + self.nb_ref_imgs), ch[0], kernel_size=7)
        self.conv2 = conv_gn(ch[0], ch[1], kernel_size=5)
        self.conv3 = conv_gn(ch[1], ch[2])
        self.conv4 = conv_gn(ch[2], ch[3])
        self.conv5 = conv_gn(ch[3], ch[4])
        self.conv6 = conv_gn(ch[4], ch[5])
        self.conv7 = conv_gn(ch[5], ch[6])
        self.pose_pred = nn.Conv2d(ch[6], 6 * self.nb_ref_imgs, kernel_size=1, padding=0)
        for m in self.modules():
            if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
                nn.init.xavier_uniform_(m.weight.data)
                if m.bias is not None:
                    m.bias.data.zero_()

    def forward(self, image, context):
        assert len(context) == self.nb_ref_imgs
        x = torch.cat([image] + list(context), 1)
        for layer in (self.conv1, self.conv2, self.conv3, self.conv4, self.conv5, self.conv6, self.conv7):
            x = layer(x)
        pose = self.pose_pred(x).mean(3).mean(2)
        return 0.01 * pose.view(pose.size(0), self.nb_ref_imgs, 6)


def require_fp32_library_convolutions():
    """PoseNet stays on the library's convolutions (north_star).  PyTorch lets cuDNN run them in TF32 by default
    (torch.backends.cudnn.allow_tf32 = True), which moves the predicted poses by ~1e-4 relative to the fp32 reference --
    measured on the B200 against tests/golden/step_2x64x96.npz -- and the photometric loss with them.  The parity bar of this
    path is fp32, so the step harness switches the library to fp32 convolutions (PoseNet is 0.2 % of the step's FLOPs)."""
    torch.backends.cudnn.allow_tf32 = False


YACS_LOSS_DEFAULTS = dict(num_scales=4, ssim_loss_weight=0.85, occ_reg_weight=0.1, smooth_loss_weight=0.001,
                          C1=1e-4, C2=9e-4, photometric_reduce_op='min', disp_norm=True, clip_loss=0.0,
                          progressive_scaling=0.0, padding_mode='zeros', automask_loss=True)
YACS_MODEL_DEFAULTS = dict(rotation_mode='euler', flip_lr_prob=0.5, upsample_depth_maps=True)


class SelfSupModel(nn.Module):
    """Self-supervised SfM model: depth net + pose net + photometric loss.

    forward(batch, progress) follows SfmModel.compute_depth_net (random left-right flip around the depth
    net, nearest up-sampling of the 4 scales, SfmModel.py:81-90), compute_pose_net (:92-96) and
    SelfSupModel.forward (:63-97).  Defaults = the yacs training defaults (configs/default_config.py:88-103)."""

    def __init__(self, depth_net=None, pose_net=None, rotation_mode='euler', flip_lr_prob=0.5,
                 upsample_depth_maps=True, fuse_upsample=True, pose_stream=True, **loss_kwargs):
        """fuse_upsample: with upsample_depth_maps, hand the four maps to the loss at their own resolution and let its kernel
        read them nearest-upsampled (losses.MultiViewPhotometricLoss.forward(nearest_upsample=True)); False = the three
        F.interpolate copies of the reference's glue.  forward() returns 'inv_depths' at the network's resolutions then."""
        super().__init__()
        require_fp32_library_convolutions()
        self.fuse_upsample = fuse_upsample
        self.pose_stream = pose_stream and os.environ.get("PN_POSE_STREAM", "1") != "0"
        self._side_stream = None
        self.depth_net = depth_net if depth_net is not None else PackNet01(version='1A')
        self.pose_net = pose_net if pose_net is not None else PoseNet(nb_ref_imgs=2, rotation_mode=rotation_mode)
        self.rotation_mode = rotation_mode
        self.flip_lr_prob = flip_lr_prob
        self.upsample_depth_maps = upsample_depth_maps
        kw = dict(YACS_LOSS_DEFAULTS)
        kw.update(loss_kwargs)
        self._photometric_loss = MultiViewPhotometricLoss(**kw)

    @property
    def logs(self):
        return dict(self._photometric_loss.logs)

    def compute_depth_net(self, rgb, force_flip=False):
        flip = (random.random() < self.flip_lr_prob) if self.training else force_flip
        out = self.depth_net(rgb=torch.flip(rgb, [3]) if flip else rgb)
        inv = out['inv_depths']
        if flip:
            inv = [torch.flip(d, [3]) for d in inv] if isinstance(inv, (list, tuple)) else torch.flip(inv, [3])
        if self.training and self.upsample_depth_maps and not self.fuse_upsample:
            shape = inv[0].shape[-2:]
            inv = [F.interpolate(d, shape, mode='nearest') for d in inv]   # model_utils.py:152-180
        return inv

    def compute_pose_net(self, image, contexts):
        pose_vec = self.pose_net(image, contexts)
        return [Pose.from_vec(pose_vec[:, i], self.rotation_mode) for i in range(pose_vec.shape[1])]

    def forward(self, batch, return_logs=False, progress=0.0):
        poses = None
        side = None
        if 'rgb_context' in batch and self.pose_stream and batch['rgb'].is_cuda:
            # PoseNet (library convolutions: 1.9 ms of small launches that leave most SMs idle) shares nothing with the depth
            # network before the loss: run it on a second stream next to PackNet01.  Autograd replays each backward op on the
            # stream of its forward, so the two backward passes overlap as well; a captured step keeps the fork / join.
            cur = torch.cuda.current_stream()
            if self._side_stream is None:
                self._side_stream = torch.cuda.Stream()
            side = self._side_stream
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                poses = self.compute_pose_net(batch['rgb'], batch['rgb_context'])
        inv_depths = self.compute_depth_net(batch['rgb'])
        if side is not None:
            torch.cuda.current_stream().wait_stream(side)
            for p in poses:
                p.mat.record_stream(torch.cuda.current_stream())
        elif 'rgb_context' in batch:
            poses = self.compute_pose_net(batch['rgb'], batch['rgb_context'])
        out = {'inv_depths': inv_depths, 'poses': poses}
        if not self.training:
            return out
        loss = self._photometric_loss(batch['rgb_original'], batch['rgb_context_original'], inv_depths,
                                      batch['intrinsics'], batch['intrinsics'], poses,
                                      return_logs=return_logs, progress=progress,
                                      nearest_upsample=self.upsample_depth_maps and self.fuse_upsample)
        return {'loss': loss['loss'], 'metrics': loss['metrics'], **out}
