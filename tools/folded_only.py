"""pack1 block forward + backward through the folded path (target of `ncu -k regex:"conv_igemm|fold_|frame_"`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from packnet_sfm_b200 import folded, functional as PF  # noqa: E402

dev = torch.device("cuda:0")
B, H, W, C, Co, k = 4, 192, 640, 64, 64, 5
x = (torch.rand(B, H, W, C, device=dev) - 0.5).requires_grad_(True)
w2 = ((torch.rand(Co, 32 * C, k, k, device=dev) - 0.5) * 0.01).requires_grad_(True)
b2 = torch.zeros(Co, device=dev, requires_grad=True)
w3 = (torch.rand(8, 1, 3, 3, 3, device=dev) - 0.5).requires_grad_(True)
b3 = torch.zeros(8, device=dev, requires_grad=True)
for _ in range(3):
    z = folded.pack_conv_folded(x, w2, b2, w3, b3, PF.conv2d)
    z.backward(torch.ones_like(z))
torch.cuda.synchronize()
print("ok", float(z.abs().mean()))
