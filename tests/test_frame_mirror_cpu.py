"""frame_kernels.cu mirror (tests/kernel_mirrors.py), driven by the SAME term table the kernel binding uses
(folded.frame_term_specs), against the PyTorch definition of the frame terms (folded.frame_strips + autograd)."""
import numpy as np
import pytest
import torch

from kernel_mirrors import border_class, frame_backward, frame_forward
from packnet_sfm_b200 import folded


def run(B, C, H, W, Co, k, seed):
    g = torch.Generator().manual_seed(seed)
    dd = dict(dtype=torch.float64, generator=g)
    n = 4 * C
    m = k // 2
    h, w = H // 2, W // 2
    xs = torch.rand(B, h, w, n, **dd) - 0.5
    lines = {"top": xs[:, 0].clone(), "bottom": xs[:, h - 1].clone(), "left": xs[:, :, 0].clone(), "right": xs[:, :, w - 1].clone()}
    for t in lines.values():
        t.requires_grad_(True)
    w2 = torch.rand(Co, 8 * n, k, k, **dd) - 0.5
    w3 = torch.rand(8, 1, 3, 3, 3, **dd) - 0.5
    b3 = torch.rand(8, **dd) - 0.5
    folds = [f.detach().clone().requires_grad_(True) for f in folded.fold_set_torch(w2, w3)]
    beta, ts, bs, ls, rs = folded.frame_strips(lines["top"], lines["bottom"], lines["left"], lines["right"], folds, b3, k)
    z_ref = torch.zeros(B, h, w, Co, dtype=torch.float64)
    z_ref[:, :m] += ts
    z_ref[:, h - m:] += bs
    z_ref[:, :, :m] += ls
    z_ref[:, :, w - m:] += rs
    gz = torch.rand(B, h, w, Co, **dd) - 0.5
    (z_ref * gz).sum().backward()
    _, dB = folded.bias_classes(folds[9].detach(), b3, k, w, "cpu")
    # ---- the mirror on flat arrays, terms from the shared table
    names = folded.FOLD_ORDER
    W_ohwi = {nm: folds[i].detach().permute(0, 2, 3, 1).contiguous().numpy().reshape(-1) for i, nm in enumerate(names) if nm != "main"}
    flat = {k_: v.detach().numpy().reshape(-1) for k_, v in lines.items()}
    dflat = {k_: np.zeros_like(v) for k_, v in flat.items()}
    terms = []
    for sp in folded.frame_term_specs(h, w, n, k):
        T = dict(sp)
        Lfull = w if sp["line"] in ("top", "bottom") else h
        T["line_bs"] = T["dline_bs"] = Lfull * n
        T["line"] = flat[sp["line"]][sp["px"] * n:]
        T["dline"] = dflat[sp["line"]][sp["px"] * n:]
        T["w"] = W_ohwi[sp["name"]]
        T["dw"] = np.zeros_like(T["w"])
        terms.append(T)
    z = np.zeros(B * h * w * Co)
    frame_forward(terms, B, h, w, Co, n, m, dB.numpy().reshape(-1), z)
    err = np.abs(z.reshape(B, h, w, Co) - z_ref.detach().numpy()).max()
    assert err < 1e-11, err
    gg = 2 * m + 1
    gdB = np.zeros(gg * gg * Co)
    frame_backward(terms, B, h, w, Co, n, m, gz.numpy().reshape(-1), gdB)
    for k_ in flat:
        e_ = np.abs(dflat[k_] - lines[k_].grad.numpy().reshape(-1)).max()
        assert e_ < 1e-11, (k_, e_)
    for T in terms:
        ref = folds[names.index(T["name"])].grad.permute(0, 2, 3, 1).contiguous().numpy().reshape(-1)
        e_ = np.abs(T["dw"] - ref).max()
        assert e_ < 1e-11, (T["name"], e_)
        assert np.abs(ref).max() > 0, T["name"]
    ref = np.zeros((gg, gg, Co))
    gzn = gz.numpy()
    for row in range(h):
        for col in range(w):
            if row < m or row >= h - m or col < m or col >= w - m:
                ref[border_class(row, h, m), border_class(col, w, m)] += gzn[:, row, col].sum(0)
    assert np.abs(gdB.reshape(gg, gg, Co) - ref).max() < 1e-11


@pytest.mark.parametrize("case", [(1, 1, 8, 10, 2, 3, 0), (2, 1, 12, 10, 3, 5, 1)])
def test_frame_kernel_mirror_matches_the_pytorch_frame_terms(case):
    run(*case)
