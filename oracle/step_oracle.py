"""CPU oracle of ONE self-supervised training step (the reference's hot loop), fp32 PyTorch:
PackNet01 forward (oracle/packnet_oracle.py) -> nearest upsample -> PoseNet -> Pose.from_vec -> photometric
loss (oracle/loss_oracle.py) -> backward -> Adam(lr 2e-4).

TEST INFRASTRUCTURE ONLY.  Used by bench.py as the `cpu_baseline` / `--impl reference` leg (kind "port": the
Python reference itself cannot travel to the GPU box) and by tests.  Restates
  packnet_sfm/trainers/horovod_trainer.py:85-96, models/SelfSupModel.py:63-97, models/SfmModel.py:81-127,
  networks/pose/PoseNet.py:38-86, models/model_wrapper.py:128-166 (Adam, lr 2e-4)."""
import math

import torch
import torch.nn.functional as F

from . import loss_oracle as LO
from . import packnet_oracle as PO


def posenet_state_dict(seed=43, nb_ref_imgs=2):
    g = torch.Generator().manual_seed(seed)
    ch = [3 * (1 + nb_ref_imgs), 16, 32, 64, 128, 256, 256, 256]
    ks = [7, 5, 3, 3, 3, 3, 3]
    sd = {}
    for i in range(7):
        k = ks[i]
        bound = math.sqrt(6.0 / ((ch[i] + ch[i + 1]) * k * k))
        sd["conv%d.0.weight" % (i + 1)] = (torch.rand(ch[i + 1], ch[i], k, k, generator=g) * 2 - 1) * bound
        sd["conv%d.0.bias" % (i + 1)] = torch.zeros(ch[i + 1])
        sd["conv%d.1.weight" % (i + 1)] = torch.ones(ch[i + 1])
        sd["conv%d.1.bias" % (i + 1)] = torch.zeros(ch[i + 1])
    bound = math.sqrt(6.0 / (256 + 6 * nb_ref_imgs))
    sd["pose_pred.weight"] = (torch.rand(6 * nb_ref_imgs, 256, 1, 1, generator=g) * 2 - 1) * bound
    sd["pose_pred.bias"] = torch.zeros(6 * nb_ref_imgs)
    return sd


def posenet_forward(image, context, sd, nb_ref_imgs=2):
    """PoseNet.forward, PoseNet.py:67-84."""
    x = torch.cat([image] + list(context), 1)
    ks = [7, 5, 3, 3, 3, 3, 3]
    for i in range(7):
        p = "conv%d" % (i + 1)
        x = F.conv2d(x, sd[p + ".0.weight"], sd[p + ".0.bias"], stride=2, padding=(ks[i] - 1) // 2)
        x = F.relu(F.group_norm(x, 16, sd[p + ".1.weight"], sd[p + ".1.bias"], 1e-5))
    pose = F.conv2d(x, sd["pose_pred.weight"], sd["pose_pred.bias"]).mean(3).mean(2)
    return 0.01 * pose.view(pose.size(0), nb_ref_imgs, 6)


class StepOracle:
    def __init__(self, depth_sd=None, pose_sd=None, lr=2e-4, device="cpu"):
        """device="cuda": the same plain-PyTorch program on the GPU's library kernels (cuDNN / ATen) -- bench.py's
        `stock_torch_gpu` leg, the number BASELINE.md asks the product to beat; never a parity reference."""
        mk = lambda v: v.clone().to(device).requires_grad_(True)      # noqa: E731
        self.depth = {k: mk(v) for k, v in (depth_sd or PO.packnet01_state_dict(42)).items()}
        self.pose = {k: mk(v) for k, v in (pose_sd or posenet_state_dict(43)).items()}
        self.opt = torch.optim.Adam(list(self.depth.values()) + list(self.pose.values()), lr=lr)

    def forward_loss(self, batch, flip=False):
        rgb = batch["rgb"]
        inv = PO.packnet01_forward(torch.flip(rgb, [3]) if flip else rgb, self.depth)
        if flip:
            inv = [torch.flip(d, [3]) for d in inv]
        inv = LO.upsample_output(inv)
        vec = posenet_forward(rgb, batch["rgb_context"], self.pose)
        poses = [LO.pose_from_vec(vec[:, i]) for i in range(vec.shape[1])]
        K = batch["intrinsics"]
        return LO.multiview_photometric_loss(batch["rgb"], batch["rgb_context"], inv, K, K, poses)

    def step(self, batch, flip=False):
        self.opt.zero_grad(set_to_none=True)
        out = self.forward_loss(batch, flip)
        out["loss"].backward()
        self.opt.step()
        return float(out["loss"].detach())
