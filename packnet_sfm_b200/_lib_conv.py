"""ctypes declarations for the convolution / layer entry points of include/packnet_b200.h."""
import ctypes

PRECISION_TF32X1 = 1
PRECISION_BF16X1 = 2
PRECISION_TF32X3 = 3
PRECISION_BF16X3 = 4
MODE_AUTO, MODE_PER_TAP, MODE_HALO = 0, 1, 2


class ConvDesc(ctypes.Structure):
    _fields_ = [
        ("batch", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("cin", ctypes.c_int32), ("cout", ctypes.c_int32), ("ksize", ctypes.c_int32),
        ("precision", ctypes.c_int32), ("mode", ctypes.c_int32), ("debug_flags", ctypes.c_int32),
    ]


class FoldDesc(ctypes.Structure):
    _fields_ = [("cout", ctypes.c_int32), ("n", ctypes.c_int32), ("ksize", ctypes.c_int32),
                ("ky0", ctypes.c_int32), ("ky1", ctypes.c_int32), ("kx0", ctypes.c_int32), ("kx1", ctypes.c_int32),
                ("dy0", ctypes.c_int32), ("dy1", ctypes.c_int32), ("dx0", ctypes.c_int32), ("dx1", ctypes.c_int32),
                ("layout", ctypes.c_int32)]


class FrameTerm(ctypes.Structure):
    _fields_ = [("line", ctypes.c_void_p), ("w", ctypes.c_void_p), ("dline", ctypes.c_void_p), ("dw", ctypes.c_void_p),
                ("line_bstride", ctypes.c_int64), ("dline_bstride", ctypes.c_int64),
                ("w_sco", ctypes.c_int64), ("w_sa", ctypes.c_int64), ("w_se", ctypes.c_int64),
                ("L", ctypes.c_int32), ("A", ctypes.c_int32), ("A2", ctypes.c_int32), ("KE", ctypes.c_int32), ("pad", ctypes.c_int32),
                ("r0", ctypes.c_int32), ("ra1", ctypes.c_int32), ("ra2", ctypes.c_int32), ("rl", ctypes.c_int32),
                ("c0", ctypes.c_int32), ("ca1", ctypes.c_int32), ("ca2", ctypes.c_int32), ("cl", ctypes.c_int32),
                ("alpha", ctypes.c_float), ("bias_mode", ctypes.c_int32)]


class FrameDesc(ctypes.Structure):
    _fields_ = [("batch", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
                ("cout", ctypes.c_int32), ("n", ctypes.c_int32), ("ksize", ctypes.c_int32), ("num_terms", ctypes.c_int32),
                ("flags", ctypes.c_int32), ("terms", FrameTerm * 8)]


def declare(lib):
    c = ctypes
    vp, sz = c.c_void_p, c.c_size_t
    lib.pn_conv2d_forward.argtypes = [c.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp, vp]
    lib.pn_conv2d_dgrad.argtypes = [c.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp]
    lib.pn_conv2d_dgrad.restype = c.c_int
    lib.pn_conv2d_rows_pad.argtypes = [c.c_int]
    lib.pn_conv2d_rows_pad.restype = c.c_int
    lib.pn_conv2d_packed_weight_elems.argtypes = [c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.POINTER(sz)]
    lib.pn_conv2d_pack_weight.argtypes = [vp, vp, vp, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, vp]
    lib.pn_tf32_residual.argtypes = [vp, vp, sz, vp]
    lib.pn_split_bf16.argtypes = [vp, vp, vp, sz, vp]
    lib.pn_split_bf16.restype = c.c_int
    i, f = c.c_int, c.c_float
    lib.pn_feature_stencil_forward.argtypes = [i, vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
    lib.pn_feature_stencil_backward.argtypes = [i, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, vp]
    lib.pn_feature_stencil_backward_parts.argtypes = [i, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, vp]
    lib.pn_groupnorm_elu_forward.argtypes = [vp, vp, vp, vp, f, vp, vp, vp, i, i, i, i, i, vp]
    lib.pn_groupnorm_elu_backward.argtypes = [vp, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, vp, vp, i, i, i, i, i, i, i, vp]
    lib.pn_groupnorm_elu_forward_split.argtypes = [vp, vp, vp, vp, f, vp, vp, vp, vp, i, i, i, vp]
    lib.pn_groupnorm_elu_backward_split.argtypes = [vp, vp, vp, vp, vp, f, vp, vp, vp, vp, vp, vp, vp, vp, i, i, i, vp]
    lib.pn_groupnorm_elu_forward_split.restype = lib.pn_groupnorm_elu_backward_split.restype = c.c_int
    lib.pn_channel_sum.argtypes = [vp, vp, sz, i, vp]
    lib.pn_head_conv_forward.argtypes = [vp, vp, vp, vp, i, i, i, i, vp]
    lib.pn_head_conv_backward.argtypes = [vp, vp, vp, vp, vp, vp, i, i, i, i, vp]
    for name in ("pn_feature_stencil_forward", "pn_feature_stencil_backward", "pn_feature_stencil_backward_parts", "pn_groupnorm_elu_forward",
                 "pn_groupnorm_elu_backward", "pn_channel_sum", "pn_head_conv_forward", "pn_head_conv_backward"):
        getattr(lib, name).restype = c.c_int
    lib.pn_conv2d_wgrad.argtypes = [c.POINTER(ConvDesc), vp, vp, vp, vp, vp, vp, vp]
    lib.pn_conv2d_unpack_weight_grad.argtypes = [vp, vp, i, i, i, i, vp]
    lib.pn_conv2d_pack_weight_tiled.argtypes = [vp, vp, vp, i, i, i, i, i, vp]
    lib.pn_conv2d_pack_weight_tiled.restype = c.c_int
    lib.pn_conv2d_unpack_weight_grad_tiled.argtypes = [vp, vp, i, i, i, i, vp]
    lib.pn_conv2d_unpack_weight_grad_tiled.restype = c.c_int
    lib.pn_conv2d_wgrad.restype = c.c_int
    lib.pn_conv2d_unpack_weight_grad.restype = c.c_int
    lib.pn_conv2d_wgrad_packed_elems.argtypes = [i, i, i, i, c.POINTER(sz)]
    lib.pn_conv2d_wgrad_packed_elems.restype = c.c_int
    for name in ("pn_conv2d_forward", "pn_conv2d_packed_weight_elems", "pn_conv2d_pack_weight", "pn_tf32_residual"):
        getattr(lib, name).restype = c.c_int
    lib.pn_pack_fold_forward.argtypes = [c.POINTER(FoldDesc), vp, vp, vp, vp]
    lib.pn_pack_fold_backward.argtypes = [c.POINTER(FoldDesc), vp, vp, vp, vp, vp, vp, i, vp]
    lib.pn_pack_frame_forward.argtypes = [c.POINTER(FrameDesc), vp, vp, vp]
    lib.pn_pack_frame_backward.argtypes = [c.POINTER(FrameDesc), vp, vp, vp]
    lib.pn_pack_frame_forward.restype = c.c_int
    lib.pn_pack_frame_backward.restype = c.c_int
    lib.pn_pack_fold_set_forward.argtypes = [i, i, i, vp, vp, c.POINTER(vp), vp]
    lib.pn_pack_fold_set_backward.argtypes = [i, i, i, vp, vp, c.POINTER(vp), vp, vp, vp, vp]
    lib.pn_pack_fold_set_forward.restype = c.c_int
    lib.pn_pack_fold_set_backward.restype = c.c_int
    lib.pn_pack_fold_forward.restype = c.c_int
    lib.pn_pack_fold_backward.restype = c.c_int
    return lib
