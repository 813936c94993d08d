"""Build-container only (marker `ref`): the oracle restatement against the LIVE unmodified reference."""
import pytest
import torch

from conftest import rel_l2

pytestmark = pytest.mark.ref


def test_packnet01_restatement_is_bit_identical_to_reference_module():
    from oracle import ref_shims
    ref_shims.install()
    from packnet_sfm.networks.depth.PackNet01 import PackNet01
    from oracle import packnet_oracle as PO
    from packnet_sfm_b200 import synthetic
    torch.manual_seed(0)
    net = PackNet01(version="1A").train()
    sd = {k: v.detach() for k, v in net.state_dict().items()}
    sd2 = PO.packnet01_state_dict()
    assert set(sd) == set(sd2) and all(sd[k].shape == sd2[k].shape for k in sd)
    x = synthetic.make_frames(1, 64, 96, seed=3)["rgb"]
    with torch.no_grad():
        ref = net(x)["inv_depths"]
        mine = PO.packnet01_forward(x, sd)
    for a, b in zip(ref, mine):
        assert torch.equal(a, b)


def test_loss_restatement_matches_reference_class():
    from oracle import ref_shims
    ref_shims.install()
    from packnet_sfm.losses.multiview_photometric_loss import MultiViewPhotometricLoss
    from packnet_sfm.geometry.pose import Pose
    from oracle import loss_oracle as LO
    from packnet_sfm_b200 import synthetic
    B, H, W = 2, 24, 48
    fr = synthetic.make_frames(B, H, W, seed=5)
    inv = synthetic.make_inv_depths(B, H, W, seed=6)
    mats = [LO.pose_from_vec(synthetic.make_pose_vecs(B, seed=7)[:, j]) for j in range(2)]
    ref = MultiViewPhotometricLoss(num_scales=4, ssim_loss_weight=0.85, smooth_loss_weight=0.001,
                                   photometric_reduce_op="min", clip_loss=0.0, automask_loss=True)
    K = fr["intrinsics"]
    a = ref(fr["rgb"], fr["rgb_context"], inv, K, K, [Pose(m.clone()) for m in mats])
    b = LO.multiview_photometric_loss(fr["rgb"], fr["rgb_context"], inv, K, K, mats)
    assert float(a["loss"]) == float(b["loss"])
