"""pack1 convolution [4,96,320,2048] x [64,2048,5,5] (target of `ncu -k regex:conv_igemm_kernel`)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from packnet_sfm_b200 import functional as PF  # noqa: E402

prec = {"x1": PF.PRECISION_TF32X1, "tf32x3": PF.PRECISION_TF32X3}.get(sys.argv[1] if len(sys.argv) > 1 else "", PF.PRECISION_BF16X3)
PF.set_precision(prec)
dev = torch.device("cuda:0")
x = torch.rand(4, 96, 320, 2048, device=dev) - 0.5
w = (torch.rand(64, 2048, 5, 5, device=dev) - 0.5) * 0.01
wp, wlo = PF._pack_weight(w, False, prec)
xh, xlo = PF._operands(x, prec)
for _ in range(3):
    y = PF._conv_raw(xh, xlo, wp, wlo, None, 64, 5, prec)
torch.cuda.synchronize()
print("ok", float(y.abs().mean()))
