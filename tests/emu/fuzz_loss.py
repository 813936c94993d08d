"""Randomised parity sweep of the fused loss kernels under the host emulation (tests/emu/) against the oracle: random
sizes (ragged tiles, maps smaller than a tile), 1..3 context frames, 1..4 scales (full-resolution and multi-resolution),
automask on/off, min / mean, large and small motion; loss to 2e-5, gradient fields (kink pixels excluded as in
tests/test_loss_gpu.py), warp tap indices bit-exact.  Sizes whose coarsest scale is under 3 pixels are refused loudly by
the library (ReflectionPad2d(1) is undefined there) and show up as EXC lines.  TEST INFRASTRUCTURE; not collected by pytest:
python tests/emu/fuzz_loss.py <seed> <seconds>.  Round 1: ~220 valid cases over three seeds, no mismatch."""
import ctypes, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, numpy as np
from conftest import rel_l2
from packnet_sfm_b200 import _lib, _lib_conv, synthetic
lib = ctypes.CDLL(os.path.join(ROOT, 'tests', 'emu', '_build', 'libpacknet_emu.so'))   # python -c 'import __graft_entry__ as g; g.build()' builds it
_lib._declare(lib); _lib_conv.declare(lib)
_lib.lib = lambda: lib; _lib.require_cuda = lambda *a: None; _lib.current_stream = lambda: None
from packnet_sfm_b200.losses import MultiViewPhotometricLoss, warp_tap_indices
from packnet_sfm_b200.geometry import Pose
from oracle import loss_oracle as LO
random.seed(int(sys.argv[1])); t_end=time.time()+float(sys.argv[2]); n=0
while time.time()<t_end:
    B=random.choice([1,1,2]); H=random.randint(4,40); W=random.randint(4,70); N=random.choice([1,2,2,3])
    full=random.random()<0.6
    if not full:
        H=(H//8+1)*8; W=(W//8+1)*8
    cfg=dict(num_scales=random.choice([1,2,3,4,4]), ssim_loss_weight=random.choice([0.85,0.5]), smooth_loss_weight=random.choice([0.001,0.0,0.1]),
             photometric_reduce_op="min", clip_loss=0.0, automask_loss=random.random()<0.7)
    if not cfg["automask_loss"] and random.random()<0.5: cfg["photometric_reduce_op"]="mean"
    seed=random.randint(0,10000)
    try:
        fr=synthetic.make_frames(B,H,W,seed=seed)
        ctx=list(fr["rgb_context"])
        while len(ctx)<N: ctx.append(torch.roll(fr["rgb"],2+len(ctx),3)*0.9+0.05)
        ctx=ctx[:N]
        inv=synthetic.make_inv_depths(B,H,W,seed=seed+1,full_res=full)
        g=torch.Generator().manual_seed(seed)
        big = random.random()<0.3
        sc=torch.tensor([0.2,0.2,0.2,0.02,0.02,0.02])*(5.0 if big else 1.0)
        mats=[LO.pose_from_vec((torch.rand(B,6,generator=g)-0.5)*sc) for _ in range(N)]
        K=fr["intrinsics"]
        inv_d=[d.clone().requires_grad_(True) for d in inv]; mats_d=[m.clone().requires_grad_(True) for m in mats]
        out=MultiViewPhotometricLoss(**cfg)(fr["rgb"],ctx,inv_d,K,K,[Pose(m) for m in mats_d]); out["loss"].backward()
        inv_c=[d.clone().requires_grad_(True) for d in inv]; mats_c=[m.clone().requires_grad_(True) for m in mats]
        ref=LO.multiview_photometric_loss(fr["rgb"],ctx,inv_c,K,K,mats_c,**{k:v for k,v in cfg.items() if k!="clip_loss"}); ref["loss"].backward()
        a,b=float(out["loss"].item()),float(ref["loss"].item())
        ok = abs(a-b)<=2e-5*abs(b)
        ge=[]
        for i in range(cfg["num_scales"]):
            ga, gb = inv_d[i].grad, inv_c[i].grad
            if ga is None or gb is None: continue
            err=(ga.double()-gb.double()).abs(); scale=float(gb.abs().max())+1e-30
            outl=(err>1e-3*scale); frac=float(outl.double().mean())
            inl=~outl
            rel=float((err[inl]**2).sum().sqrt()/((gb.double()[inl]**2).sum().sqrt()+1e-30))
            ge.append((round(frac,4), rel))
            if frac>max(2e-3, (21.0 if os.environ.get('PN_LOSS_GROUPED')=='1' else 12.0)/err.numel()) or rel>1e-3: ok=False
        # warp indices bit exact at scale 0 context 0 (only meaningful full-res)
        if full:
            taps,coords=warp_tap_indices(inv[0],K,K,mats[0]); idx,oc=LO.warp_tap_indices(inv[0],K,K,mats[0])
            if not (np.array_equal(coords.numpy().view(np.int32), oc.view(np.int32)) and np.array_equal(taps.numpy(), idx)): ok=False; ge.append("WARPIDX")
        n+=1
        if not ok: print("MISMATCH", (B,H,W,N,full,big,seed), cfg, a, b, ge, flush=True)
    except Exception as ex:
        print("EXC", (B,H,W,N,full,seed), cfg, repr(ex)[:300], flush=True)
print("cases", n)
