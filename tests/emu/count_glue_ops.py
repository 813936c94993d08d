"""(glue-op census: see the bottom)  One whole self-supervised step (PackNet01 + PoseNet + photometric loss, forward AND backward) on the CPU through the
product's own Python layer, twice: with the default kernels and with the staged variants switched on (grouped-scale loss program, im2col first layer, flat tile staging, GroupNorm tree statistics; `fold` adds the folded pack layers) -- all SIMT kernels run from
their real source under the host emulation (tests/emu/), only the tcgen05 convolution is PyTorch's conv2d.  The two runs
must agree on the loss and on every parameter gradient: the integration check of `bench.py --staged-all` that does not need
a GPU.  TEST INFRASTRUCTURE; a script: python tests/emu/step_emulated.py [fold]   (about four minutes; with `fold` the
emulated fold / frame backward kernels take it to roughly half an hour)"""
import ctypes, inspect, os, random, sys, textwrap, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch
import torch.nn.functional as F
from packnet_sfm_b200 import _lib, _lib_conv, folded, functional as PF, losses, synthetic

lib = ctypes.CDLL(os.path.join(ROOT, 'tests', 'emu', '_build', 'libpacknet_emu.so'))   # __graft_entry__.build() builds it
_lib._declare(lib); _lib_conv.declare(lib)
_lib.lib = lambda: lib; _lib.require_cuda = lambda *a: None; _lib.current_stream = lambda: None
folded._use_kernels = lambda t: True


def conv2d(x, w, b=None):
    """stand-in for the tcgen05 engine (differentiable): NHWC in / out, zero pad k//2, weight zero-padded to the input's channels"""
    if w.shape[1] != x.shape[3]:
        w = torch.cat([w, torch.zeros(w.shape[0], x.shape[3] - w.shape[1], w.shape[2], w.shape[3])], 1)
    return F.conv2d(x.permute(0, 3, 1, 2), w, b, padding=w.shape[-1] // 2).permute(0, 2, 3, 1).contiguous()


PF.conv2d = conv2d
from packnet_sfm_b200.networks import PackNet01   # noqa: E402
from packnet_sfm_b200.models import SelfSupModel   # noqa: E402

src = textwrap.dedent(inspect.getsource(PackNet01.forward)).replace("if not rgb.is_cuda:", "if False:")
ns = {}
exec("import torch\nimport torch.nn.functional as F\nfrom packnet_sfm_b200 import _lib, functional as PF\n"
     "from packnet_sfm_b200.networks import _cat_channels\n" + src, ns)
PackNet01.forward = ns["forward"]


WITH_FOLD = "fold" in sys.argv[1:]


def run(staged):
    PF.set_pack_fold(staged and WITH_FOLD, min_pixels=0)
    PF.set_im2col_first(staged)
    losses.set_grouped_kernel(staged)
    _lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, int(staged))
    _lib.set_tuning(_lib.PN_TUNE_GN_TREE, int(staged))
    torch.manual_seed(42); random.seed(3)
    model = SelfSupModel(flip_lr_prob=0.0).train()
    B, H, W = 1, 64, 96
    fr = synthetic.make_frames(B, H, W, seed=11)
    batch = {"rgb": fr["rgb"], "rgb_context": fr["rgb_context"], "rgb_original": fr["rgb"],
             "rgb_context_original": fr["rgb_context"], "intrinsics": fr["intrinsics"]}
    t0 = time.time()
    out = model(batch)
    out["loss"].backward()
    print("%s step: %.1f s, loss %.8f" % ("staged" if staged else "default", time.time() - t0, float(out["loss"])))
    return float(out["loss"]), {k: p.grad.clone() for k, p in model.named_parameters() if p.grad is not None}



# ---- census of the PyTorch (ATen) operations a step issues AROUND the library kernels: what the "ATen glue" lines of the step
# breakdown consist of.   python tests/emu/count_glue_ops.py
import collections
from torch.profiler import profile, ProfilerActivity
PF.set_pack_fold(False); PF.set_im2col_first(True); losses.set_grouped_kernel(True)
_lib.set_tuning(_lib.PN_TUNE_STAGE_FLAT, 1); _lib.set_tuning(_lib.PN_TUNE_GN_TREE, 1)
torch.manual_seed(42); random.seed(3)
model = SelfSupModel(flip_lr_prob=0.0).train()
B, H, W = 1, 64, 96
fr = synthetic.make_frames(B, H, W, seed=11)
batch = {"rgb": fr["rgb"], "rgb_context": fr["rgb_context"], "rgb_original": fr["rgb"],
         "rgb_context_original": fr["rgb_context"], "intrinsics": fr["intrinsics"]}
out = model(batch)
with profile(activities=[ProfilerActivity.CPU], record_shapes=True) as prof:
    out["loss"].backward()
cnt = collections.Counter()
for e in prof.events():
    if e.name in ("aten::add", "aten::add_", "aten::copy_", "aten::mul", "aten::cat", "aten::fill_", "aten::zero_", "aten::sum", "aten::contiguous", "aten::clone"):
        cnt[(e.name, str(e.input_shapes)[:90])] += 1
print("backward: ATen elementwise / copy operations by input shapes")
for (n, sh), c in sorted(cnt.items(), key=lambda kv: (kv[0][0], -kv[1])):
    print("%4d  %-16s %s" % (c, n, sh))
