"""Generate tests/golden/*.npz from the LIVE, UNMODIFIED reference (build container only).

TEST INFRASTRUCTURE ONLY.  Run:  python oracle/gen_golden.py
/root/reference cannot travel to the GPU box, so the vectors it produces are committed as small
fixtures together with this script (the reference has no golden vectors of its own).

Every output array comes from reference classes:
  packnet_sfm.losses.multiview_photometric_loss.MultiViewPhotometricLoss  (+ autograd for the gradients)
  packnet_sfm.geometry.camera.Camera.reconstruct / project                 (normalised warp coordinates)
  packnet_sfm.networks.layers.packnet.layers01.{PackLayerConv3d,UnpackLayerConv3d,Conv2D,ResidualConv}
  packnet_sfm.networks.depth.PackNet01.PackNet01
Inputs and weights are regenerated from seeds (packnet_sfm_b200/synthetic.py, oracle/packnet_oracle.py);
the inputs are stored too so an RNG change cannot silently invalidate a fixture.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

ref_shims.install()
from oracle import packnet_oracle as PO  # noqa: E402
from packnet_sfm_b200 import synthetic  # noqa: E402
from oracle.loss_oracle import pose_from_vec, upsample_output  # noqa: E402

from packnet_sfm.losses.multiview_photometric_loss import MultiViewPhotometricLoss  # noqa: E402
from packnet_sfm.geometry.camera import Camera  # noqa: E402
from packnet_sfm.geometry.pose import Pose  # noqa: E402
from packnet_sfm.networks.layers.packnet import layers01 as L  # noqa: E402
from packnet_sfm.networks.depth.PackNet01 import PackNet01  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
YACS_LOSS_DEFAULTS = dict(num_scales=4, ssim_loss_weight=0.85, occ_reg_weight=0.1, smooth_loss_weight=0.001,
                          C1=1e-4, C2=9e-4, photometric_reduce_op="min", disp_norm=True, clip_loss=0.0,
                          progressive_scaling=0.0, padding_mode="zeros", automask_loss=True)


def loss_case(name, B, H, W, full_res, **overrides):
    frames = synthetic.make_frames(B, H, W, seed=11)
    inv = synthetic.make_inv_depths(B, H, W, seed=12, full_res=full_res)
    vec = synthetic.make_pose_vecs(B, seed=13)
    if overrides.pop("big_motion", False):
        vec = vec * 8.0  # push many samples out of bounds / behind the camera
    inv = [d.clone().requires_grad_(True) for d in inv]
    mats = [pose_from_vec(vec[:, j]).clone().requires_grad_(True) for j in range(vec.shape[1])]
    poses = [Pose(m) for m in mats]
    kw = dict(YACS_LOSS_DEFAULTS)
    kw.update(overrides)
    loss_fn = MultiViewPhotometricLoss(**kw)
    K = frames["intrinsics"]
    out = loss_fn(frames["rgb"], frames["rgb_context"], inv, K, K, poses)
    out["loss"].backward()
    arrays = {
        "rgb": frames["rgb"], "ctx0": frames["rgb_context"][0], "ctx1": frames["rgb_context"][1],
        "K": K, "loss": out["loss"].detach(),
        "photometric_loss": out["metrics"]["photometric_loss"], "smoothness_loss": out["metrics"]["smoothness_loss"],
    }
    for i, d in enumerate(inv):
        arrays["inv%d" % i] = d.detach()
        arrays["ginv%d" % i] = d.grad
    for j, m in enumerate(mats):
        arrays["pose%d" % j] = m.detach()
        arrays["gpose%d" % j] = m.grad
    # normalised warp coordinates straight from the reference camera classes (scale 0)
    with torch.no_grad():
        for j in range(2):
            cam = Camera(K=K.float())
            ref_cam = Camera(K=K.float(), Tcw=Pose(mats[j].detach()))
            depth = 1.0 / inv[0].detach().clamp(min=1e-6)
            arrays["grid%d" % j] = ref_cam.project(cam.reconstruct(depth, frame="w"), frame="w")
    meta = {k: v for k, v in kw.items()}
    np.savez_compressed(os.path.join(OUT, name + ".npz"), meta=np.array(repr(meta)),
                        **{k: np.asarray(v.detach().numpy() if torch.is_tensor(v) else v) for k, v in arrays.items()})
    print(name, float(out["loss"]))


def _load(module, sd):
    missing = module.load_state_dict(sd, strict=True)
    return module


def block_cases():
    arrays = {}
    g = torch.Generator().manual_seed(5)

    def run(tag, module, sd, x):
        _load(module, sd)
        x = x.clone().requires_grad_(True)
        y = module(x)
        gy = torch.rand(y.shape, generator=g) - 0.5
        y.backward(gy)
        arrays[tag + "_x"] = x.detach()
        arrays[tag + "_y"] = y.detach()
        arrays[tag + "_gy"] = gy
        arrays[tag + "_gx"] = x.grad
        for k, p in module.named_parameters():
            arrays[tag + "_g_" + k] = p.grad

    run("pack_k3", L.PackLayerConv3d(32, 3), PO.block_state_dict("pack", 32, k=3, seed=21),
        torch.rand(2, 32, 16, 24, generator=g) - 0.5)
    run("pack_k5", L.PackLayerConv3d(16, 5), PO.block_state_dict("pack", 16, k=5, seed=22),
        torch.rand(1, 16, 12, 40, generator=g) - 0.5)
    run("unpack", L.UnpackLayerConv3d(64, 32, 3), PO.block_state_dict("unpack", 64, 32, 3, seed=23),
        torch.rand(2, 64, 6, 20, generator=g) - 0.5)
    run("conv2d_k7", L.Conv2D(32, 32, 7, 1), PO.block_state_dict("conv2d", 32, 32, 7, seed=24),
        torch.rand(1, 32, 12, 20, generator=g) - 0.5)
    run("residual", L.ResidualConv(32, 64, 1), PO.block_state_dict("residual", 32, 64, seed=25),
        torch.rand(2, 32, 8, 12, generator=g) - 0.5)
    np.savez_compressed(os.path.join(OUT, "blocks.npz"), **{k: v.numpy() for k, v in arrays.items()})
    print("blocks", len(arrays))


def packnet_case():
    sd = PO.packnet01_state_dict(seed=42, randomize_affine=True)
    net = PackNet01(version="1A")
    _load(net, sd)
    net.train()
    x = synthetic.make_frames(1, 64, 96, seed=31)["rgb"]
    with torch.no_grad():
        out = net(x)["inv_depths"]
    arrays = {"rgb": x.numpy()}
    for i, d in enumerate(out):
        arrays["disp%d" % (i + 1)] = d.numpy()
    np.savez_compressed(os.path.join(OUT, "packnet01_64x96.npz"), **arrays)
    print("packnet01", [tuple(d.shape) for d in out], float(out[0].mean()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    loss_case("loss_fullres", 2, 32, 64, True)
    loss_case("loss_multires", 2, 32, 64, False)
    loss_case("loss_mean_noautomask", 1, 24, 40, True, photometric_reduce_op="mean", automask_loss=False,
              smooth_loss_weight=0.1)
    loss_case("loss_bigmotion", 2, 32, 64, True, big_motion=True)
    loss_case("loss_progressive", 1, 32, 64, False, progressive_scaling=0.2)
    block_cases()
    packnet_case()
