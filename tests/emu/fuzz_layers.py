"""Randomised parity sweep of the layer kernels (feature stencils, GroupNorm+ELU, head convolution) under the host
emulation (tests/emu/): random shapes -- odd channel counts, maps smaller than a tile, ragged widths -- against float64
PyTorch.  TEST INFRASTRUCTURE; not collected by pytest (minutes of CPU):  python tests/emu/fuzz_layers.py <seed> <seconds>
Round 1: 289 cases, no mismatch."""
import ctypes, os, random, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import torch, torch.nn.functional as F
from conftest import rel_l2
from packnet_sfm_b200 import _lib, _lib_conv
from oracle import packnet_oracle as PO
lib = ctypes.CDLL(os.path.join(ROOT, 'tests', 'emu', '_build', 'libpacknet_emu.so'))   # python -c 'import __graft_entry__ as g; g.build()' builds it
_lib._declare(lib); _lib_conv.declare(lib)
_lib.lib = lambda: lib; _lib.require_cuda = lambda *a: None; _lib.current_stream = lambda: None
from packnet_sfm_b200 import functional as PF
random.seed(int(sys.argv[1]) if len(sys.argv)>1 else 0)
torch.manual_seed(0)
t_end = time.time() + float(sys.argv[2] if len(sys.argv)>2 else 120)
n=0
while time.time() < t_end:
    kind = random.choice(["pack","unpack","gn","head"])
    B = random.choice([1,1,2,3])
    if kind == "pack":
        h, w = random.randint(1,9), random.randint(1,20); C = random.choice([1,2,3,4,8,16,32,64,6,10])
        shape = (B, 2*h, 2*w, C)
    elif kind == "unpack":
        h, w = random.randint(1,9), random.randint(1,20); C = random.choice([4,8,16,24,32,64,12,20])
        shape = (B, h, w, C)
    elif kind == "gn":
        h, w = random.randint(1,9), random.randint(1,20); C = random.choice([16,32,64,128,48,80])
        shape = (B, h, w, C)
    else:
        h, w = random.randint(1,20), random.randint(1,40); C = random.choice([4,8,16,32,64,12])
        shape = (B, h, w, C)
    try:
        x = (torch.rand(*shape) - 0.5).requires_grad_(True)
        if kind in ("pack","unpack"):
            w3 = (torch.rand(8,1,3,3,3)-0.5).requires_grad_(True); b3=(torch.rand(8)-0.5).requires_grad_(True)
            y = PF.pack_features(x,w3,b3) if kind=="pack" else PF.unpack_features(x,w3,b3)
            gy = torch.rand_like(y)-0.5; y.backward(gy)
            xd,wd,bd = (t.detach().double().requires_grad_(True) for t in (x,w3,b3))
            xn = xd.permute(0,3,1,2)
            yr = (PO.conv3d_features(PO.packing(xn),wd,bd) if kind=="pack" else F.pixel_shuffle(PO.conv3d_features(xn,wd,bd),2)).permute(0,2,3,1)
            yr.backward(gy.double())
            errs = [rel_l2(y,yr), rel_l2(x.grad,xd.grad), rel_l2(w3.grad,wd.grad), rel_l2(b3.grad,bd.grad)]
            tol = [1e-6,1e-5,1e-4,1e-4]
        elif kind == "gn":
            g=(torch.rand(shape[3])+0.5).requires_grad_(True); bt=(torch.rand(shape[3])-0.5).requires_grad_(True)
            use2 = random.random()<0.5
            x2=(torch.rand(*shape)-0.5).requires_grad_(True)
            y = PF.groupnorm_elu(x*2-0.2,g,bt,1e-5,x2=x2 if use2 else None)
            gy=torch.rand_like(y)-0.5; y.backward(gy)
            xd,x2d,gd,bd=(t.detach().double().requires_grad_(True) for t in (x,x2,g,bt))
            inp = xd*2-0.2 + (x2d if use2 else 0)
            yr=F.elu(F.group_norm(inp.permute(0,3,1,2),16,gd,bd,1e-5)).permute(0,2,3,1); yr.backward(gy.double())
            errs=[rel_l2(y,yr), rel_l2(x.grad,xd.grad), rel_l2(g.grad,gd.grad), rel_l2(bt.grad,bd.grad)]; tol=[1e-5,1e-4,1e-4,1e-4]
        else:
            wt=((torch.rand(1,shape[3],3,3)-0.5)*0.2).requires_grad_(True); b=(torch.rand(1)-0.5).requires_grad_(True)
            y=PF.head_conv(x,wt,b); gy=torch.rand_like(y)-0.5; y.backward(gy)
            xd,wd,bd=(t.detach().double().requires_grad_(True) for t in (x,wt,b))
            yr=F.conv2d(xd.permute(0,3,1,2),wd,bd,padding=1)[:,0]; yr.backward(gy.double())
            errs=[rel_l2(y,yr), rel_l2(x.grad,xd.grad), rel_l2(wt.grad,wd.grad), rel_l2(b.grad,bd.grad)]; tol=[1e-6,1e-6,1e-5,1e-5]
        bad = [ (e,t) for e,t in zip(errs,tol) if not (e<t)]
        n+=1
        if bad: print("MISMATCH", kind, shape, errs, flush=True)
    except Exception as ex:
        print("EXC", kind, shape, repr(ex)[:200], flush=True)
print("cases", n)
