"""Python mirror of packnet_sfm_b200/csrc/frame_kernels.cu -- the loaders, the (a, l) -> (row, col) maps, the weight
strides and the frame-pixel enumeration of the forward / line-gradient / weight-gradient / bias-class kernels -- driven by
the SAME term table the kernel binding uses (folded.frame_term_specs) and checked against the PyTorch definition of the
frame terms (folded.frame_strips + autograd).  The kernels themselves only run on the GPU tier; this pins their index
arithmetic on the CPU tier.  Keep in step with the .cu file."""
import numpy as np
import pytest
import torch
from packnet_sfm_b200 import folded

def border_class(p, ln, m): return p if p < m else (p - (ln - m) + m + 1 if p >= ln - m else m)

def zmap(T, a, l):
    a1, a2 = a // T["A2"], a % T["A2"]
    return T["r0"] + a1*T["ra1"] + a2*T["ra2"] + l*T["rl"], T["c0"] + a1*T["ca1"] + a2*T["ca2"] + l*T["cl"]

def widx(T, co, a, e, nn): return co*T["w_sco"] + a*T["w_sa"] + e*T["w_se"] + nn

def load_line(T, line, b, l, e, nn):
    # line: [B, Lfull, n] full border line; T.px selects the start pixel; T.L the length
    pos = l + e - T["pad"]
    if pos < 0 or pos >= T["L"]: return 0.0
    return line[b, T["px"] + pos, nn]

def run(B, C, H, W, Co, k, seed):
    g = torch.Generator().manual_seed(seed)
    dd = dict(dtype=torch.float64, generator=g)
    n = 4*C; m = k//2; h, w = H//2, W//2
    xs = torch.rand(B, h, w, n, **dd) - 0.5
    lines = {"top": xs[:, 0].clone(), "bottom": xs[:, h-1].clone(), "left": xs[:, :, 0].clone(), "right": xs[:, :, w-1].clone()}
    for t in lines.values(): t.requires_grad_(True)
    w2 = ((torch.rand(Co, 8*n, k, k, **dd) - 0.5)).requires_grad_(True)
    w3 = (torch.rand(8, 1, 3, 3, 3, **dd) - 0.5).requires_grad_(True)
    b3 = (torch.rand(8, **dd) - 0.5).requires_grad_(True)
    folds = folded.fold_set_torch(w2, w3)
    folds_leaf = [f.detach().clone().requires_grad_(True) for f in folds]
    beta, ts, bs, ls, rs = folded.frame_strips(lines["top"], lines["bottom"], lines["left"], lines["right"], folds_leaf, b3, k)
    z_ref = torch.zeros(B, h, w, Co, dtype=torch.float64)
    z_ref[:, :m] += ts; z_ref[:, h-m:] += bs; z_ref[:, :, :m] += ls; z_ref[:, :, w-m:] += rs
    gz = torch.rand(B, h, w, Co, **dd) - 0.5
    _, dB = folded.bias_classes(folds_leaf[9], b3, k, w, "cpu")
    dB_leaf = dB.detach().clone().requires_grad_(True)
    # reference grads wrt lines, folded weights (OIHW), dB: rebuild z_ref with dB as a leaf
    beta2, ts, bs, ls, rs = folded.frame_strips(lines["top"], lines["bottom"], lines["left"], lines["right"], folds_leaf, b3, k)
    (z_ref * gz).sum().backward()
    # ---- mirror
    specs = folded.frame_term_specs(h, w, n, k)
    names = folded.FOLD_ORDER
    W_ohwi = {nm: folds[i].detach().permute(0, 2, 3, 1).contiguous().numpy().reshape(-1) for i, nm in enumerate(names) if nm != "main"}
    Ln = {k_: v.detach().numpy() for k_, v in lines.items()}
    dBn = dB.detach().numpy()
    z = np.zeros((B, h, w, Co))
    for T in specs:
        Wv = W_ohwi[T["name"]]; line = Ln[T["line"]]
        M, N, K = B*T["L"], T["A"]*Co, T["KE"]*n
        for p in range(M):
            b, l = p // T["L"], p % T["L"]
            for c in range(N):
                a, co = c // Co, c % Co
                acc = 0.0
                for kk in range(K):
                    e, nn = kk // n, kk % n
                    acc += load_line(T, line, b, l, e, nn) * Wv[widx(T, co, a, e, nn)]
                row, col = zmap(T, a, l)
                v = T["alpha"]*acc
                if T["bias_mode"] == 1 or (T["bias_mode"] == 2 and m <= row < h - m):
                    v += dBn[border_class(row, h, m), border_class(col, w, m), co]
                z[b, row, col, co] += v
    err = np.abs(z - z_ref.detach().numpy()).max()
    assert err < 1e-11, err
    # backward line
    gzn = gz.numpy()
    dl = {k_: np.zeros_like(v) for k_, v in Ln.items()}
    dws = {}
    for T in specs:
        Wv = W_ohwi[T["name"]]; line = Ln[T["line"]]
        M, N, K = B*T["L"], n, T["A"]*T["KE"]*Co
        for p in range(M):
            b, j = p // T["L"], p % T["L"]
            for nn in range(N):
                acc = 0.0
                for kk in range(K):
                    co, ae = kk % Co, kk // Co
                    e, a = ae % T["KE"], ae // T["KE"]
                    l = j - e + T["pad"]
                    if l < 0 or l >= T["L"]: continue
                    row, col = zmap(T, a, l)
                    acc += gzn[b, row, col, co] * Wv[widx(T, co, a, e, nn)]
                dl[T["line"]][b, T["px"] + j, nn] += T["alpha"]*acc
        # backward weight
        Mw, Nw, Kw = T["A"]*Co, T["KE"]*n, B*T["L"]
        dw = np.zeros_like(Wv)
        for r in range(Mw):
            a, co = r // Co, r % Co
            for c in range(Nw):
                e, nn = c // n, c % n
                acc = 0.0
                for p in range(Kw):
                    b, l = p // T["L"], p % T["L"]
                    row, col = zmap(T, a, l)
                    acc += gzn[b, row, col, co] * load_line(T, line, b, l, e, nn)
                dw[widx(T, co, a, e, nn)] = T["alpha"]*acc
        dws[T["name"]] = dw
    for k_ in dl:
        e_ = np.abs(dl[k_] - lines[k_].grad.numpy()).max(); assert e_ < 1e-11, (k_, e_)
    for i, nm in enumerate(names):
        if nm == "main": continue
        ref = folds_leaf[i].grad.permute(0, 2, 3, 1).contiguous().numpy().reshape(-1)
        e_ = np.abs(dws[nm] - ref).max(); assert e_ < 1e-11, (nm, e_)
        assert np.abs(ref).max() > 0, nm
    # backward bias classes
    F_ = h*w - (h-2*m)*(w-2*m); gg = 2*m+1
    gdB = np.zeros((gg, gg, Co))
    for b in range(B):
        for f in range(F_):
            if f < m*w: row, col = f // w, f % w
            elif f < 2*m*w: q = f - m*w; row, col = h - m + q // w, q % w
            else:
                q = f - 2*m*w; row = m + q // (2*m); cc = q % (2*m); col = cc if cc < m else w - 2*m + cc
            gdB[border_class(row, h, m), border_class(col, w, m)] += gzn[b, row, col]
    # reference: d/d(dB) of sum(z*gz) where z frame pixels get dB[class]: compute directly
    ref = np.zeros_like(gdB)
    for row in range(h):
        for col in range(w):
            if row < m or row >= h-m or col < m or col >= w-m:
                ref[border_class(row, h, m), border_class(col, w, m)] += gzn[:, row, col].sum(0)
    e_ = np.abs(gdB - ref).max(); assert e_ < 1e-11, e_
    # and the forward's bias placement is covered by fwd err (dB inside z_ref)



@pytest.mark.parametrize("case", [(1, 1, 8, 10, 2, 3, 0), (2, 1, 12, 10, 3, 5, 1)])
def test_frame_kernel_mirror_matches_the_pytorch_frame_terms(case):
    run(*case)
