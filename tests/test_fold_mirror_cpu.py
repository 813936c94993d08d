"""fold_fwd_kernel / fold_bwd_kernel mirror (tests/kernel_mirrors.py) against the PyTorch definition of the folds
(folded._ct3 / _perm_n): all window shapes, both output layouts, dS and the accumulate flag."""
import numpy as np
import torch

from kernel_mirrors import bwd, fwd
from packnet_sfm_b200 import folded

import pytest


@pytest.mark.parametrize("cout,C,k,ohwi", [(2, 3, 3, 0), (1, 33, 3, 1), (2, 2, 5, 1), (1, 2, 5, 0)])
def test_fold_kernel_mirror_matches_the_pytorch_folds(cout, C, k, ohwi):
    torch.manual_seed(0)
    n = 4*C; m = k//2
    w2 = torch.rand(cout, 8*n, k, k, dtype=torch.float64)-0.5
    w3 = torch.rand(8,1,3,3,3, dtype=torch.float64)-0.5
    w2r = w2.reshape(cout, 8, n, k, k)
    lo, hi = (0, m), (m+1, k)
    wins = {"main": ((0,k),(0,k),(0,3),(0,3)), "top": (lo,(0,k),(2,3),(0,3)), "bot": (hi,(0,k),(0,1),(0,3)),
            "left": ((0,k),lo,(0,3),(2,3)), "right": ((0,k),hi,(0,3),(0,1)), "tl": (lo,lo,(2,3),(2,3)), "br": (hi,hi,(0,1),(0,1)),
            "tr": (lo,hi,(2,3),(0,1))}
    dw2_k = np.zeros(w2.numel()); dw3_k = np.zeros(216)
    w2g = w2.clone().requires_grad_(True); w3g = w3.clone().requires_grad_(True)
    total = 0
    first = True
    for name, (ky, kx, dy, dx) in wins.items():
        ref = folded._perm_n(folded._ct3(w2r[:, :, :, ky[0]:ky[1], kx[0]:kx[1]], w3[:, :, :, dy[0]:dy[1], dx[0]:dx[1]]))
        got = fwd(w2.numpy(), w3.numpy(), cout, n, k, ky[0], ky[1], kx[0], kx[1], dy[0], dy[1], dx[0], dx[1], ohwi)
        if ohwi:
            got = got.transpose(0, 3, 1, 2)
        err = np.abs(got - ref.numpy()).max()
        assert not np.isnan(got).any(), name
        # backward
        g = torch.rand_like(ref)-0.5
        refg = folded._perm_n(folded._ct3(w2g.reshape(cout,8,n,k,k)[:, :, :, ky[0]:ky[1], kx[0]:kx[1]], w3g[:, :, :, dy[0]:dy[1], dx[0]:dx[1]]))
        total = total + (refg*g).sum()
        dS = (torch.rand(cout,8,k,k,dtype=torch.float64)-0.5) if first else None
        if dS is not None:
            total = total + (w2g.reshape(cout,8,n,k,k).sum(2)*dS).sum()
        gk = g.permute(0, 2, 3, 1).contiguous().numpy() if ohwi else g.numpy()
        bwd(w2.numpy(), w3.numpy(), gk, dS.numpy() if dS is not None else None, dw2_k, dw3_k, 0 if first else 1,
            cout, n, k, ky[0], ky[1], kx[0], kx[1], dy[0], dy[1], dx[0], dx[1], ohwi)
        first = False
        assert err < 1e-12, (name, err)
    total.backward()
    e2 = np.abs(dw2_k - w2g.grad.reshape(-1).numpy()).max(); e3 = np.abs(dw3_k - w3g.grad.reshape(-1).numpy()).max()
    assert e2 < 1e-12 and e3 < 1e-12

