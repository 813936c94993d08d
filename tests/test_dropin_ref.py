"""Build container only: the drop-in registration makes the UNMODIFIED reference glue resolve to our classes."""
import sys

import pytest

pytestmark = pytest.mark.ref


def test_reference_models_pick_up_the_dropin_classes():
    from oracle import ref_shims
    ref_shims.install()
    for name in list(sys.modules):
        if name.startswith("packnet_sfm.models") or name in ("packnet_sfm.losses.multiview_photometric_loss",
                                                             "packnet_sfm.networks.depth.PackNet01"):
            del sys.modules[name]
    from packnet_sfm_b200 import dropin, losses, networks
    dropin.install()
    try:
        from packnet_sfm.models.SelfSupModel import SelfSupModel        # reference file, unmodified
        from packnet_sfm.utils.load import load_class                   # reference plug-in mechanism
        model = SelfSupModel(num_scales=4, photometric_reduce_op="min", automask_loss=True, clip_loss=0.0)
        assert type(model._photometric_loss) is losses.MultiViewPhotometricLoss
        cls = load_class("PackNet01", paths=["packnet_sfm.networks.depth"])
        assert cls is networks.PackNet01
        net = cls(version="1A", dropout=0.0)
        assert len(net.state_dict()) == 216
    finally:
        ref_shims.install()     # drops the aliases again so later tests see the real reference
        for name in list(sys.modules):
            if name.startswith("packnet_sfm.models"):
                del sys.modules[name]


def _purge(prefixes):
    for name in list(sys.modules):
        if any(name == p or name.startswith(p + ".") for p in prefixes):
            del sys.modules[name]


def test_reference_selfsup_forward_runs_unmodified_through_the_dropin(emulated_kernels, monkeypatch):
    """the drop-in must EXECUTE, not only resolve.  The reference's own SelfSupModel.forward (SfmModel flip /
    upsample_output / PoseNet / Pose.from_vec glue, unmodified files) is run twice on the same seeded batch and weights: with
    the reference's PackNet01 + MultiViewPhotometricLoss, and after dropin.install() with this package's classes -- whose SIMT
    kernels (feature stencils, GroupNorm+ELU, head convolution, fused loss) execute from their real source under the host
    emulation; only the tcgen05 convolution is stood in by PyTorch's fp32 conv2d.  Same loss, same metrics, same depth maps."""
    import torch
    import torch.nn.functional as F
    from oracle import packnet_oracle as PO, ref_shims
    from oracle.step_oracle import posenet_state_dict
    from packnet_sfm_b200 import dropin, functional as PF, networks, synthetic, losses
    H, W = 32, 64
    fr = synthetic.make_frames(1, H, W, seed=3)
    batch = {"rgb": fr["rgb"], "rgb_context": fr["rgb_context"], "rgb_original": fr["rgb"],
             "rgb_context_original": fr["rgb_context"], "intrinsics": fr["intrinsics"]}
    cfg = dict(num_scales=4, ssim_loss_weight=0.85, smooth_loss_weight=0.001, photometric_reduce_op="min", clip_loss=0.0,
               automask_loss=True, rotation_mode="euler", flip_lr_prob=1.0, upsample_depth_maps=True)   # the flip path, forced
    depth_sd, pose_sd = PO.packnet01_state_dict(seed=42, randomize_affine=True), posenet_state_dict(43)

    def run():
        from packnet_sfm.models.SelfSupModel import SelfSupModel          # the reference's file, whatever it resolves to
        from packnet_sfm.networks.pose.PoseNet import PoseNet
        from packnet_sfm.utils.load import load_class
        depth = load_class("PackNet01", paths=["packnet_sfm.networks.depth"])(version="1A")
        depth.load_state_dict(depth_sd, strict=True)
        pose = PoseNet(nb_ref_imgs=2, rotation_mode="euler")
        pose.load_state_dict(pose_sd, strict=True)
        model = SelfSupModel(depth_net=depth, pose_net=pose, **cfg).train()
        with torch.no_grad():
            out = model(batch)
        return model, out

    ref_shims.install()
    _purge(["packnet_sfm.models", "packnet_sfm.losses.multiview_photometric_loss", "packnet_sfm.networks.depth.PackNet01",
            "packnet_sfm.networks.layers.packnet.layers01"])
    ref_model, ref_out = run()
    assert type(ref_model.depth_net).__module__.startswith("packnet_sfm.")

    def conv2d(x, w, b=None):       # stand-in for the tensor-core engine (NHWC, zero pad k//2, weight padded to the activation)
        if w.shape[1] != x.shape[3]:
            w = torch.cat([w, torch.zeros(w.shape[0], x.shape[3] - w.shape[1], w.shape[2], w.shape[3])], 1)
        return F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=w.shape[-1] // 2).permute(0, 2, 3, 1).contiguous()

    monkeypatch.setattr(PF, "conv2d", conv2d)
    _purge(["packnet_sfm.models"])
    dropin.install()
    try:
        model, out = run()
        assert type(model.depth_net) is networks.PackNet01 and type(model._photometric_loss) is losses.MultiViewPhotometricLoss
        want, got = float(ref_out["loss"]), float(out["loss"])
        assert abs(got - want) <= 1e-5 * abs(want), (got, want)
        for key in ("photometric_loss", "smoothness_loss"):
            assert abs(float(out["metrics"][key]) - float(ref_out["metrics"][key])) <= 1e-4 * abs(float(ref_out["metrics"][key])) + 1e-9
        for a, b in zip(out["inv_depths"], ref_out["inv_depths"]):
            assert a.shape == b.shape and float(((a - b).abs() / b.abs()).max()) < 1e-4
    finally:
        ref_shims.install()
        _purge(["packnet_sfm.models"])
