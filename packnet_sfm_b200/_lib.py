"""ctypes loader for libpacknet_b200.so (the C-ABI in include/packnet_b200.h).

There is no CPU fallback: if the library is missing or a call fails, the op raises."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libpacknet_b200.so")
PN_MAX_SCALES = 4
PN_MAX_CONTEXT = 4

_lock = threading.Lock()
_lib = None


class LossDesc(ctypes.Structure):
    _fields_ = [
        ("batch", ctypes.c_int32), ("height", ctypes.c_int32), ("width", ctypes.c_int32),
        ("num_context", ctypes.c_int32), ("num_scales", ctypes.c_int32),
        ("scale_h", ctypes.c_int32 * PN_MAX_SCALES), ("scale_w", ctypes.c_int32 * PN_MAX_SCALES),
        ("ssim_loss_weight", ctypes.c_float), ("smooth_loss_weight", ctypes.c_float),
        ("C1", ctypes.c_float), ("C2", ctypes.c_float),
        ("reduce_min", ctypes.c_int32), ("automask", ctypes.c_int32), ("flags", ctypes.c_int32),
        ("inv_shift", ctypes.c_int32 * PN_MAX_SCALES),
    ]


PN_LOSS_FLAG_GROUPED = 1
PN_TUNE_STAGE_FLAT = 1
PN_TUNE_GN_TREE = 2




def _declare(lib):
    c = ctypes
    vp, sz, fp = c.c_void_p, c.c_size_t, c.c_void_p
    lib.pn_version.restype = c.c_int
    lib.pn_last_error_string.restype = c.c_char_p
    lib.pn_launch_count.restype = c.c_uint64
    lib.pn_set_tuning.argtypes = [c.c_int, c.c_int]
    lib.pn_set_tuning.restype = c.c_int
    lib.pn_loss_workspace_bytes.argtypes = [c.POINTER(LossDesc), c.POINTER(sz)]
    lib.pn_loss_forward.argtypes = [c.POINTER(LossDesc), fp, c.POINTER(vp), c.POINTER(vp), fp, fp, c.POINTER(vp), fp,
                                    vp, sz, vp]
    lib.pn_loss_backward.argtypes = [c.POINTER(LossDesc), fp, c.POINTER(vp), c.POINTER(vp), fp, fp, c.POINTER(vp), fp,
                                     c.POINTER(vp), c.POINTER(vp), vp, sz, vp]
    lib.pn_loss_forward_backward.argtypes = [c.POINTER(LossDesc), fp, c.POINTER(vp), c.POINTER(vp), fp, fp, c.POINTER(vp), fp,
                                             c.POINTER(vp), c.POINTER(vp), sz, vp, sz, vp]
    lib.pn_loss_backward_finish.argtypes = [c.POINTER(LossDesc), fp, c.POINTER(vp), c.POINTER(vp), c.POINTER(vp), c.POINTER(vp),
                                            vp, sz, vp]
    lib.pn_loss_warp_indices.argtypes = [c.POINTER(LossDesc), c.c_int, fp, fp, fp, fp, vp, vp, vp, sz, vp]
    lib.pn_resize_bilinear_ac.argtypes = [fp, fp, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, c.c_int, vp]
    for name in ("pn_loss_workspace_bytes", "pn_loss_forward", "pn_loss_backward", "pn_loss_forward_backward",
                 "pn_loss_backward_finish", "pn_loss_warp_indices",
                 "pn_resize_bilinear_ac"):
        getattr(lib, name).restype = c.c_int
    return lib


def lib():
    """The loaded library; raises RuntimeError (never falls back) when it is absent."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise RuntimeError(
                        "libpacknet_b200.so is not built (%s). Run `python -m packnet_sfm_b200.build` "
                        "or __graft_entry__.build(); there is no CPU/PyTorch fallback." % LIB_PATH)
                loaded = ctypes.CDLL(LIB_PATH)
                _declare(loaded)
                from . import _lib_conv
                _lib_conv.declare(loaded)
                _lib = loaded
    return _lib


def check(rc, what):
    if rc != 0:
        msg = lib().pn_last_error_string().decode("utf-8", "replace")
        raise RuntimeError("%s failed (code %d): %s" % (what, rc, msg))


def set_tuning(key, value):
    """Process-wide switch of a STAGED kernel variant (include/packnet_b200.h pn_set_tuning)."""
    check(lib().pn_set_tuning(int(key), int(value)), "pn_set_tuning")


def launch_count():
    return int(lib().pn_launch_count())


def ptr(t):
    return ctypes.c_void_p(t.data_ptr())


def ptr_array(tensors):
    arr = (ctypes.c_void_p * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr


def current_stream(device=None):
    """The caller's stream ON THE TENSORS' DEVICE (the library only enqueues on the stream it is handed)."""
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def require_cuda(*tensors):
    """Every operand must be a CUDA tensor on the CURRENT device (one process per GPU: the wrapper never switches devices
    and the kernels are enqueued on that device's current stream); a tensor of another GPU would be reached from the wrong
    context: raise instead."""
    import torch
    cur = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("packnet_sfm_b200 ops run on CUDA tensors only (got %s); there is no CPU path"
                               % t.device)
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise RuntimeError("packnet_sfm_b200 ops run on the current CUDA device (cuda:%d); got a tensor on %s -- call "
                               "torch.cuda.set_device() first (one process per GPU)" % (cur, t.device))


def require_f32(*tensors):
    """require_cuda + every operand fp32.  A model cast with .half() / .double() (the reference's `--half` switches) or an
    autocast output would otherwise be read through the raw pointer as 4-byte floats: out-of-bounds garbage, silently."""
    import torch
    require_cuda(*tensors)
    for t in tensors:
        if t is not None and t.dtype != torch.float32:
            raise RuntimeError("packnet_sfm_b200 kernels are fp32 (got %s): cast the module / inputs back with .float(); "
                               "half / double / autocast tensors are not reinterpreted silently" % t.dtype)
