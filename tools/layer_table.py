"""Per-call device time of every convolution / stencil entry point in one training step (pn_trace_*)."""
import collections
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from packnet_sfm_b200 import _lib, parallel  # noqa: E402
from packnet_sfm_b200.models import SelfSupModel  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(42)
model = SelfSupModel().to(dev).train()
from packnet_sfm_b200 import optim  # noqa: E402
from packnet_sfm_b200.networks import native_conv_weights  # noqa: E402
opt = optim.FlatAdam(model.parameters(), lr=2e-4, native=native_conv_weights(model.depth_net, (192, 640)))
batch = bench.to_device(bench.make_host_batch(4, 192, 640, 0), dev)
lib = _lib.lib()
lib.pn_trace_enable.argtypes = [ctypes.c_int]
lib.pn_trace_dump.argtypes = [ctypes.c_char_p, ctypes.c_size_t]
lib.pn_trace_dump.restype = ctypes.c_int


def step():
    opt.zero_grad()
    out = model(batch)
    out["loss"].backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
lib.pn_trace_enable(1)
step()
torch.cuda.synchronize()
lib.pn_trace_enable(0)
buf = ctypes.create_string_buffer(1 << 20)
n = lib.pn_trace_dump(buf, len(buf))
agg = collections.OrderedDict()
for line in buf.value.decode().strip().split("\n"):
    tag, ms = line.rsplit("\t", 1)
    a = agg.setdefault(tag, [0, 0.0])
    a[0] += 1
    a[1] += float(ms)
rows = sorted(agg.items(), key=lambda kv: -kv[1][1])
tot = sum(v[1] for v in agg.values())
print("%d traced calls, %.2f ms (event time around each call: includes its memsets / multi-launch sequences)" % (n, tot))
for tag, (cnt, ms) in rows:
    extra = ""
    p = dict((k.rstrip("0123456789"), int(k[len(k.rstrip("0123456789")):])) for k in tag.split()[1:] if k[-1].isdigit())
    if tag.startswith("conv_"):
        fl = 2.0 * p["B"] * p["H"] * p["W"] * p["Cin"] * p["Cout"] * p["k"] ** 2 * cnt
        extra = "%7.1f TFLOP/s" % (fl / ms / 1e9)
    print("%8.3f ms %5.1f%% x%-2d %-60s %s" % (ms, 100 * ms / tot, cnt, tag, extra))
