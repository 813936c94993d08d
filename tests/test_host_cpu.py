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


def test_pose_shim_matches_oracle():
    """The closed-form R = Rx·Ry·Rz of geometry.rigid_from_vec against the oracle's restatement of pose_utils.py:8-51
    (values and the gradient that reaches PoseNet)."""
    from packnet_sfm_b200.geometry import Pose
    from oracle import loss_oracle as LO
    vec = (torch.rand(3, 6) - 0.5).requires_grad_(True)
    ours, ref = Pose.from_vec(vec, "euler").mat, LO.pose_from_vec(vec)
    assert torch.allclose(ours, ref, atol=1e-7, rtol=0)
    w = torch.rand(3, 4, 4)
    g1, = torch.autograd.grad((ours * w).sum(), vec)
    g2, = torch.autograd.grad((ref * w).sum(), vec)
    assert torch.allclose(g1, g2, atol=1e-6, rtol=0)
    assert len(Pose.identity(2)) == 2 and torch.equal(Pose.identity(2).mat[1], torch.eye(4))
    with pytest.raises(ValueError):
        Pose.from_vec(vec, "quat")


def test_fold_and_frame_descriptors_reach_the_c_side_intact():
    """The ctypes mirrors of pn_fold_desc / pn_frame_desc against the C structs: the library's own host-side validation
    (which runs before any launch) echoes the fields it read in its error text."""
    from packnet_sfm_b200 import _lib, folded
    from packnet_sfm_b200._lib_conv import FoldDesc
    lib = _lib.lib()
    buf = torch.zeros(64)
    p = _lib.ptr(buf)
    # fold: an unsupported window (2 taps of a 3x3 kernel) must be reported with exactly these numbers
    d = FoldDesc(7, 40, 3, 0, 2, 0, 3, 2, 3, 0, 3, 1)
    assert lib.pn_pack_fold_forward(ctypes.byref(d), p, p, p, None) == -2
    assert b"window 2x3 with face 1x3" in lib.pn_last_error_string(), lib.pn_last_error_string()
    d = FoldDesc(7, 40, 3, 0, 1, 0, 3, 2, 3, 0, 3, 5)
    assert lib.pn_pack_fold_forward(ctypes.byref(d), p, p, p, None) == -1
    assert b"layout 5" in lib.pn_last_error_string(), lib.pn_last_error_string()
    # frame: the real descriptor builder on CPU tensors; corrupt one field of one term and read it back from the error
    B, h, w, co, n, k = 2, 9, 11, 4, 8, 5
    lines = {"top": torch.zeros(B, w, n), "bottom": torch.zeros(B, w, n), "left": torch.zeros(B, h, n), "right": torch.zeros(B, h, n)}
    shapes = {"top": (co, 2, 7, n), "bottom": (co, 2, 7, n), "left": (co, 7, 2, n), "right": (co, 7, 2, n),
              "tl": (co, 2, 2, n), "tr": (co, 2, 2, n), "bl": (co, 2, 2, n), "br": (co, 2, 2, n)}
    weights = {name: torch.zeros(s) for name, s in shapes.items()}
    fd = folded._FrameApplyCUDA._desc((B, h, w, co), n, k, lines, weights)
    assert fd.num_terms == 8 and fd.terms[3].L == h and fd.terms[0].w_sa == 7 * n and fd.terms[2].w_se == 2 * n
    fd.terms[6].c0 = w + 5                                        # term 6 = "bl": col = c0 - a2
    z = torch.zeros(B, h, w, co)
    dB = torch.zeros(5, 5, co)
    rc = lib.pn_pack_frame_forward(ctypes.byref(fd), _lib.ptr(dB), _lib.ptr(z), None)
    msg = lib.pn_last_error_string()
    assert rc == -1 and b"term 6 maps (a=0, l=0) to (8, 16) outside the 9x11 map" in msg, (rc, msg)
    # an intact descriptor passes the validation of all eight terms; without a GPU the launch itself then fails (> 0)
    fd = folded._FrameApplyCUDA._desc((B, h, w, co), n, k, lines, weights)
    if not torch.cuda.is_available():
        assert lib.pn_pack_frame_forward(ctypes.byref(fd), _lib.ptr(dB), _lib.ptr(z), None) > 0


def test_half_and_double_tensors_are_rejected_not_reinterpreted(monkeypatch):
    """the reference's `--half` switches cast model and images; the kernels read raw fp32 pointers, so any other
    dtype must raise before a pointer is taken (the device check is bypassed here so that the dtype check is what fires)."""
    from packnet_sfm_b200 import _lib, functional as PF
    from packnet_sfm_b200.losses import MultiViewPhotometricLoss
    from packnet_sfm_b200.geometry import Pose
    monkeypatch.setattr(_lib, "require_cuda", lambda *a: None)
    x = torch.rand(1, 8, 8, 16)
    w = torch.rand(16, 16, 3, 3)
    for bad in (torch.float16, torch.float64, torch.bfloat16):
        with pytest.raises(RuntimeError, match="fp32"):
            PF.conv2d(x.to(bad), w.to(bad), None)
        with pytest.raises(RuntimeError, match="fp32"):
            PF.groupnorm_elu(x.to(bad), torch.ones(16, dtype=bad), torch.zeros(16, dtype=bad))
    loss = MultiViewPhotometricLoss(num_scales=1, photometric_reduce_op="min", clip_loss=0.0, automask_loss=True)
    img = torch.rand(1, 3, 8, 16)
    with pytest.raises(RuntimeError, match="fp32"):
        loss(img.half(), [img.half(), img.half()], [torch.rand(1, 1, 8, 16).half()], torch.eye(3)[None], torch.eye(3)[None],
             [Pose.identity(1), Pose.identity(1)])
