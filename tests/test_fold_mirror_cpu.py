"""Line-by-line Python mirror of fold_fwd_kernel / fold_bwd_kernel (packnet_sfm_b200/csrc/fold_kernels.cu): the thread ->
element maps, the shared-memory staging and the store indices of both layouts, checked against the PyTorch definition
of the folds (folded.fold_set_torch).  The kernels themselves only run on the GPU tier (tests/test_folded_gpu.py); this
keeps their index arithmetic pinned on the CPU tier.  Keep in step with the .cu file."""
import numpy as np
import torch
from packnet_sfm_b200 import folded

T = 128
def perm_channel(cpp, C): return (cpp & 3) * C + (cpp >> 2)

def fwd(w2, w3, cout, n, k, ky0, ky1, kx0, kx1, dy0, dy1, dx0, dx1, ohwi=0):
    KA, KB, DA, DB = ky1-ky0, kx1-kx0, dy1-dy0, dx1-dx0
    EA, EB = KA+DA-1, KB+DB-1; E = EA*EB; C = n//4
    w2f, w3f = w2.reshape(-1), w3.reshape(-1)
    out = np.full(cout*n*E, np.nan)
    nblk = (n+T-1)//T
    for co in range(cout):
        for bx in range(nblk):
            w3s = np.zeros(8*3*DA*DB)
            for i in range(8*3*DA*DB):
                dx = i % DB; dy = (i//DB) % DA; fdc = i//(DA*DB)
                w3s[i] = w3f[fdc*9 + (dy0+dy)*3 + (dx0+dx)]
            stage = np.zeros(T*E)
            c0 = bx*T
            for tid in range(T):
                cpp = c0+tid
                acc = np.zeros((EA, EB))
                if cpp < n:
                    for f in range(8):
                        for dc in range(3):
                            cp = cpp-dc+1
                            if cp < 0 or cp >= n: continue
                            src = ((((co*8+f)*n+cp)*k+ky0)*k+kx0)
                            v = np.array([[w2f[src+a*k+b] for b in range(KB)] for a in range(KA)])
                            wf = (f*3+dc)*DA*DB
                            for dy in range(DA):
                                for dx in range(DB):
                                    w = w3s[wf+dy*DB+dx]
                                    for a in range(KA):
                                        for b in range(KB):
                                            acc[a+dy][b+dx] += v[a][b]*w
                q = tid & 3; cl = tid >> 2
                for a in range(EA):
                    for b in range(EB):
                        stage[(q*32+cl)*E + a*EB+b] = acc[a][b]
            cbase = c0 >> 2
            nc = min(C-cbase, 32)
            for idx in range(4*32*E):
                if not ohwi:
                    qq = idx//(32*E); r = idx - qq*32*E
                    if r < nc*E:
                        out[(co*n + qq*C + cbase)*E + r] = stage[idx]
                else:
                    c = idx & 31; qq = (idx >> 5) & 3; e = idx >> 7
                    if c < nc:
                        out[(co*E + e)*n + qq*C + cbase + c] = stage[(qq*32 + c)*E + e]
    return out.reshape(cout, EA, EB, n) if ohwi else out.reshape(cout, n, EA, EB)

def bwd(w2, w3, dout, dS, dw2, dw3, accumulate, cout, n, k, ky0, ky1, kx0, kx1, dy0, dy1, dx0, dx1, ohwi=0):
    KA, KB, DA, DB = ky1-ky0, kx1-kx0, dy1-dy0, dx1-dx0
    EA, EB = KA+DA-1, KB+DB-1; E = EA*EB; C = n//4
    w2f, w3f, df = w2.reshape(-1), w3.reshape(-1), dout.reshape(-1)
    dw2f, dw3f = dw2.reshape(-1), dw3.reshape(-1)
    dSf = dS.reshape(-1) if dS is not None else None
    nblk = (n+T-1)//T
    for co in range(cout):
        for f in range(8):
            for bx in range(nblk):
                w3s = np.zeros(3*DA*DB); red = np.zeros(3*DA*DB)
                for tid in range(3*DA*DB):
                    dx = tid % DB; dy = (tid//DB) % DA; dc = tid//(DA*DB)
                    w3s[tid] = w3f[(f*3+dc)*9 + (dy0+dy)*3 + (dx0+dx)]
                for tid in range(T):
                    cp = bx*T+tid
                    if cp >= n: continue
                    g2 = np.zeros((KA,KB)); g3 = np.zeros((3,DA,DB))
                    off = ((((co*8+f)*n+cp)*k+ky0)*k+kx0)
                    v = np.array([[w2f[off+a*k+b] for b in range(KB)] for a in range(KA)])
                    for dc in range(3):
                        cpp = cp+dc-1
                        if cpp < 0 or cpp >= n: continue
                        pc = perm_channel(cpp, C)
                        d = co*E*n + pc if ohwi else (co*n + pc)*E
                        es = n if ohwi else 1
                        for ea in range(EA):
                            for eb in range(EB):
                                dv = df[d+(ea*EB+eb)*es]
                                for dy in range(DA):
                                    for dx in range(DB):
                                        a = ea-dy; b = eb-dx
                                        if 0 <= a < KA and 0 <= b < KB:
                                            g2[a][b] += dv*w3s[(dc*DA+dy)*DB+dx]
                                            g3[dc][dy][dx] += dv*v[a][b]
                    for a in range(KA):
                        for b in range(KB):
                            r = g2[a][b]
                            if dSf is not None: r += dSf[((co*8+f)*k+ky0)*k+kx0 + a*k+b]
                            if accumulate: r += dw2f[off+a*k+b]
                            dw2f[off+a*k+b] = r
                    red += g3.reshape(-1)
                for tid in range(3*DA*DB):
                    dx = tid % DB; dy = (tid//DB) % DA; dc = tid//(DA*DB)
                    dw3f[(f*3+dc)*9+(dy0+dy)*3+(dx0+dx)] += red[tid]

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

