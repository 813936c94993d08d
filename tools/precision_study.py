"""CPU study of operand-precision schemes for the tensor-core convolutions (no GPU needed): PackNet01 forward through the
oracle restatement with every Conv2d's operands quantised the way a given MMA scheme would read them, fp32 accumulation,
depth maps compared with the reference's golden vectors (tests/golden/packnet01_64x96.npz, metric of
tests/test_packnet_gpu.py: max relative error of the inverse-depth maps, bar 1e-3).

Schemes (x = activation operand, w = weight operand; "pair" = hi + lo of the 16-bit type, i.e. the value a two-term split
represents; the dropped lo*lo product of a three-product split is modelled by using both pairs):
  tf32x1      : both operands truncated to tf32 (one kind::tf32 MMA)                 -- measured on the B200: 1.3e-3 .. 8.1e-3
  bf16x3      : bf16 pairs on both sides (three kind::f16 MMAs)                       -- measured on the B200: 2.4e-5 .. 1.3e-4
  f16x3       : fp16 pairs on both sides (three MMAs)
  f16x2_w     : x = ONE fp16 (round to nearest), w = fp16 pair (two MMAs: x_hi*[w_hi; w_lo] -- one N=2bn instruction)
  f16x2_x     : x = fp16 pair, w = one fp16
  f16x1       : one fp16 on both sides (one MMA)
  bf16x2_w    : x = one bf16, w = bf16 pair
python tools/precision_study.py [--full]   (--full adds a 192x640 image against the fp32 oracle)"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import load_golden  # noqa: E402
from oracle import packnet_oracle as PO  # noqa: E402


def one(x, dt):
    return x.to(dt).to(torch.float32)


def pair(x, dt):
    hi = x.to(dt).to(torch.float32)
    return hi + (x - hi).to(dt).to(torch.float32)


def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


H, B = torch.float16, torch.bfloat16
SCHEMES = {
    "fp32": (lambda x: x, lambda w: w, 0),
    "tf32x1": (tf32_trunc, tf32_trunc, 1),
    "bf16x3": (lambda x: pair(x, B), lambda w: pair(w, B), 3),
    "f16x3": (lambda x: pair(x, H), lambda w: pair(w, H), 3),
    "f16x2_w": (lambda x: one(x, H), lambda w: pair(w, H), 2),
    "f16x2_x": (lambda x: pair(x, H), lambda w: one(w, H), 2),
    "f16x1": (lambda x: one(x, H), lambda w: one(w, H), 1),
    "bf16x2_w": (lambda x: one(x, B), lambda w: pair(w, B), 2),
}


def run(rgb, sd, scheme):
    qx, qw, _ = SCHEMES[scheme]
    stats = {"max_abs_x": 0.0, "min_nonzero_x": float("inf")}

    def conv(x, w, b, precision):
        stats["max_abs_x"] = max(stats["max_abs_x"], float(x.abs().max()))
        return F.conv2d(qx(x), qw(w), b)

    old = PO._conv2d
    PO._conv2d = conv
    try:
        with torch.no_grad():
            out = PO.packnet01_forward(rgb, sd)
    finally:
        PO._conv2d = old
    return out, stats


def main():
    torch.set_num_threads(os.cpu_count() or 8)
    z = load_golden("packnet01_64x96")
    sd = PO.packnet01_state_dict(seed=42, randomize_affine=True)
    cases = [("golden 64x96 (live reference)", z["rgb"], [z["disp%d" % i] for i in range(1, 5)])]
    if "--full" in sys.argv:
        from packnet_sfm_b200 import synthetic
        rgb = synthetic.make_frames(1, 192, 640, seed=9)["rgb"]
        ref, _ = run(rgb, sd, "fp32")
        cases.append(("synthetic 192x640 (fp32 oracle)", rgb, ref))
    for title, rgb, refs in cases:
        print(title)
        for name in SCHEMES:
            out, st = run(rgb, sd, name)
            rels = [float(((d - r).abs() / r.abs()).max()) for d, r in zip(out, refs)]
            print("  %-9s MMAs/product %d  depth max-rel per scale: %s   (max |activation| %.1f)" % (
                name, SCHEMES[name][2], " ".join("%.2e" % v for v in rels), st["max_abs_x"]))


if __name__ == "__main__":
    main()
