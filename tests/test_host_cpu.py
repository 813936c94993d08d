"""CPU tier: host logic and the C-ABI surface (no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "packnet_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(pn_[a-z0-9_]+)\s*\(", text)))


def test_library_loads_and_exports_every_declared_symbol():
    from packnet_sfm_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        from packnet_sfm_b200 import build
        build.build_library()
    lib = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert "pn_loss_forward" in syms and "pn_version" in syms
    for s in syms:
        assert hasattr(lib, s), "missing export %s" % s
    lib.pn_version.restype = ctypes.c_int
    assert lib.pn_version() >= 100


def test_descriptor_validation_errors_without_gpu():
    from packnet_sfm_b200 import _lib
    lib = _lib.lib()
    d = _lib.LossDesc()
    n = ctypes.c_size_t(0)
    assert lib.pn_loss_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) < 0
    assert b"pn_loss" in lib.pn_last_error_string()
    d.batch, d.height, d.width, d.num_context, d.num_scales = 2, 32, 64, 2, 2
    d.scale_h[0], d.scale_w[0], d.scale_h[1], d.scale_w[1] = 32, 64, 16, 32
    d.ssim_loss_weight, d.reduce_min, d.automask = 0.85, 1, 1
    assert lib.pn_loss_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) == 0
    assert n.value > 2 * 3 * 16 * 32 * 4 * 3      # room for the resized scale-1 images
    d.reduce_min = 0
    assert lib.pn_loss_workspace_bytes(ctypes.byref(d), ctypes.byref(n)) < 0   # automask needs 'min'


def test_loss_class_surface_and_no_cpu_fallback():
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss, ProgressiveScaling
    from packnet_sfm_b200.geometry import Pose
    loss = MultiViewPhotometricLoss(num_scales=4, ssim_loss_weight=0.85, occ_reg_weight=0.1, smooth_loss_weight=0.001,
                                    C1=1e-4, C2=9e-4, photometric_reduce_op="min", disp_norm=True, clip_loss=0.0,
                                    progressive_scaling=0.0, padding_mode="zeros", automask_loss=True,
                                    some_unrelated_config_key=3)
    assert loss.logs == {"num_scales": 4}
    ps = ProgressiveScaling(0.2, 4)
    assert [ps(p) for p in (0.0, 0.25, 0.45, 0.65, 1.0)] == [4, 3, 2, 1, 1]
    x = torch.rand(1, 3, 8, 16)
    with pytest.raises(RuntimeError, match="CUDA"):
        loss(x, [x, x], [torch.rand(1, 1, 8, 16)] * 4, torch.eye(3)[None], torch.eye(3)[None],
             [Pose.identity(1), Pose.identity(1)])


def test_pose_mirror_matches_oracle():
    from packnet_sfm_b200.geometry import Pose
    from oracle import loss_oracle as LO
    vec = torch.rand(3, 6) - 0.5
    assert torch.equal(Pose.from_vec(vec, "euler").mat, LO.pose_from_vec(vec))
    p = Pose.from_vec(vec, "euler")
    ident = (p @ p.inverse()).mat
    assert torch.allclose(ident, torch.eye(4).repeat(3, 1, 1), atol=1e-6)
