"""Whole-step CUDA graph (bench.py --graph): capture forward + loss + backward + Adam once, replay, and land on the same
parameters and losses as eager execution from the same state.  The capture-safety audit of the library (no allocation /
synchronisation / host copies inside the entry points) is in DESIGN.md."""
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def _setup(seed):
    from packnet_sfm_b200 import synthetic
    from packnet_sfm_b200.models import SelfSupModel
    torch.manual_seed(seed)
    model = SelfSupModel(flip_lr_prob=0.0).to(DEV).train()
    opt = torch.optim.Adam(model.parameters(), lr=2e-4, fused=True, capturable=True)
    fr = synthetic.make_frames(1, 64, 96, seed=5)
    batch = {"rgb": fr["rgb"].to(DEV), "rgb_context": [c.to(DEV) for c in fr["rgb_context"]], "intrinsics": fr["intrinsics"].to(DEV)}
    batch["rgb_original"], batch["rgb_context_original"] = batch["rgb"], batch["rgb_context"]
    return model, opt, batch


def _zero(model):
    for p in model.parameters():
        p.grad = None


def _snapshot(model, opt):
    st = []
    for group in opt.param_groups:
        for p in group["params"]:
            ps = opt.state[p]
            st.append({k: v.detach().clone() for k, v in ps.items() if torch.is_tensor(v)})
    return [p.detach().clone() for p in model.parameters()], st


def _restore(model, opt, snap):
    """in place: the captured graph holds the addresses of these tensors"""
    params, st = snap
    with torch.no_grad():
        for p, q in zip(model.parameters(), params):
            p.copy_(q)
        i = 0
        for group in opt.param_groups:
            for p in group["params"]:
                for k, v in st[i].items():
                    opt.state[p][k].copy_(v)
                i += 1


@pytest.mark.parametrize("fold", [True, False], ids=["pack_fold", "unfolded"])
def test_graph_replay_matches_eager_steps(fold):
    """Capture forward + loss + backward + Adam once, then run the SAME two steps from the SAME parameter / optimizer state once
    by replay and once eagerly.  Step 1 starts from identical parameters: its loss may differ only by the order of the
    forward's split-K atomics.  Step 2 sits one Adam update apart: elements whose gradient is zero up to rounding move by
    +-lr with the sign of the noise in both modes, so losses agree to ~1e-3 and parameters up to a few lr on a few elements."""
    from packnet_sfm_b200 import functional as PF
    PF.set_pack_fold(fold, min_pixels=0)
    try:
        model, opt, batch = _setup(3)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(2):                              # warm-up off the default stream (AccumulateGrad stream binding)
                _zero(model)
                out = model(batch)
                out["loss"].backward()
                opt.step()
            torch.cuda.current_stream().wait_stream(side)
            _zero(model)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                out = model(batch)
                out["loss"].backward()
                opt.step()
            torch.cuda.synchronize()
            snap = _snapshot(model, opt)                    # the capture itself executes nothing
            graph_losses = []
            for _ in range(2):
                g.replay()
                graph_losses.append(float(out["loss"].item()))
            graph_params = [p.detach().clone() for p in model.parameters()]
            _restore(model, opt, snap)
            eager_losses = []
            for _ in range(2):
                _zero(model)
                o = model(batch)
                o["loss"].backward()
                opt.step()
                eager_losses.append(float(o["loss"].item()))
            torch.cuda.synchronize()
        print("graph", graph_losses, "eager", eager_losses)
        assert abs(graph_losses[0] - eager_losses[0]) <= 2e-6 * abs(eager_losses[0]), (graph_losses, eager_losses)
        assert abs(graph_losses[1] - eager_losses[1]) <= 2e-3 * abs(eager_losses[1]), (graph_losses, eager_losses)
        close = total = 0
        for p, q in zip(model.parameters(), graph_params):
            d = (p - q).abs()
            assert float(d.max()) <= 2 * 2 * 2e-4 + 1e-6        # two steps, at most +-lr each, in both runs
            close += int((d <= 1e-5 + 1e-3 * q.abs()).sum())
            total += d.numel()
        assert close >= 0.999 * total, (close, total)
    finally:
        PF.set_pack_fold(True, min_pixels=1920)     # the default policy
