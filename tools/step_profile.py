"""Per-kernel device-time breakdown of one training step (CUPTI via torch.profiler; no ncu serialisation)."""
import argparse
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from packnet_sfm_b200 import functional as PF, parallel  # noqa: E402
from packnet_sfm_b200.models import SelfSupModel  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--height", type=int, default=192)
ap.add_argument("--width", type=int, default=640)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--torch-adam", action="store_true", help="torch.optim.Adam + per-call weight packing (round-1 optimizer path)")
a = ap.parse_args()

dev = torch.device("cuda:0")
torch.manual_seed(42)
model = SelfSupModel().to(dev).train()
if a.torch_adam:
    bucket = parallel.FlatBucket(model.parameters())
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, fused=True)
    zero_grad = bucket.zero_grad
else:
    from packnet_sfm_b200 import optim
    from packnet_sfm_b200.networks import native_conv_weights
    opt = optim.FlatAdam(model.parameters(), lr=2e-4, native=native_conv_weights(model.depth_net, (a.height, a.width)))
    zero_grad = opt.zero_grad
batch = bench.to_device(bench.make_host_batch(a.batch, a.height, a.width, 0), dev)


def step():
    zero_grad()
    out = model(batch)
    out["loss"].backward()
    opt.step()


for _ in range(3):
    step()
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(a.steps):
        step()
    torch.cuda.synchronize()
rows = []
for e in prof.key_averages():
    t = getattr(e, "device_time_total", None)
    if t is None:
        t = getattr(e, "cuda_time_total", 0.0)
    if t > 0:
        rows.append((t / a.steps / 1e3, e.count / a.steps, e.key))
rows.sort(reverse=True)
total = sum(r[0] for r in rows)
print("device time per step: %.2f ms over %d kernel names" % (total, len(rows)))
for ms, n, k in rows[:a.top]:
    print("%8.3f ms %5.1f%% %7.1f launches  %s" % (ms, 100 * ms / total, n, k[:110]))
