"""Whole-step parity: two optimizer steps of the self-supervised model (PackNet01 + PoseNet + loss + backward + Adam, the flip
path forced off then on) against tests/golden/step_2x64x96.npz, produced by the LIVE reference's own SelfSupModel /
PackNet01 / PoseNet / MultiViewPhotometricLoss with torch.optim.Adam (oracle/gen_golden.py::step_case;
models/SelfSupModel.py:63-97, SfmModel.py:81-127, model_wrapper.py:128-166)."""
import pytest
import torch

from conftest import load_golden, rel_l2
from oracle import packnet_oracle as PO
from oracle.step_oracle import posenet_state_dict

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def strided_index(numel, n):
    n = min(n, numel)
    return (torch.arange(n, dtype=torch.int64) * (numel - 1)) // max(n - 1, 1)


def _model():
    from packnet_sfm_b200.models import SelfSupModel
    model = SelfSupModel(flip_lr_prob=0.0)
    model.depth_net.load_state_dict(PO.packnet01_state_dict(seed=42, randomize_affine=True), strict=True)
    model.pose_net.load_state_dict(posenet_state_dict(43), strict=True)
    return model.to(DEV).train()


def _batch(z):
    from packnet_sfm_b200 import synthetic
    fr = synthetic.make_frames(int(z["B"]), int(z["H"]), int(z["W"]), seed=int(z["seed_frames"]))
    b = {"rgb": fr["rgb"].to(DEV), "rgb_context": [c.to(DEV) for c in fr["rgb_context"]], "intrinsics": fr["intrinsics"].to(DEV)}
    b["rgb_original"], b["rgb_context_original"] = b["rgb"], b["rgb_context"]
    return b


@pytest.mark.parametrize("optimizer", ["flat_native", "flat_plain", "torch"])
def test_two_training_steps_match_the_reference(optimizer):
    z = load_golden("step_2x64x96")
    model, batch = _model(), _batch(z)
    from packnet_sfm_b200 import optim
    if optimizer == "torch":
        opt = torch.optim.Adam([{"params": model.depth_net.parameters(), "lr": 2e-4},
                                {"params": model.pose_net.parameters(), "lr": 2e-4}])
    else:
        from packnet_sfm_b200.networks import native_conv_weights
        # flat_native: the convolution weights stored in the engine's layout (tiles written by the optimizer launch, weight
        # gradients accumulated in place, data gradients from the forward tiles); flat_plain: same optimizer, OIHW storage
        native = native_conv_weights(model.depth_net, (int(z["H"]), int(z["W"]))) if optimizer == "flat_native" else ()
        opt = optim.FlatAdam([{"params": list(model.depth_net.parameters())}, {"params": list(model.pose_net.parameters())}],
                             lr=2e-4, native=native)
        assert (len(native) > 40) == (optimizer == "flat_native")
    for step, flip in enumerate((0.0, 1.0)):
        model.flip_lr_prob = flip
        opt.zero_grad(set_to_none=True)
        out = model(batch)
        out["loss"].backward()
        loss = float(out["loss"].item())
        want = float(z["loss%d" % step])
        print("step %d (flip %.0f): loss %.7f reference %.7f rel %.2e" % (step, flip, loss, want, abs(loss - want) / want))
        assert abs(loss - want) <= 1e-3 * want
        for key in ("photometric_loss", "smoothness_loss"):
            got, ref = float(out["metrics"][key]), float(z[key + str(step)])
            assert abs(got - ref) <= 1e-3 * abs(ref), (key, got, ref)
        if step == 0:
            assert ((out["inv_depths"][0].detach().cpu() - z["inv_depth0_step0"]).abs() / z["inv_depth0_step0"].abs()).max() < 1e-3
            assert out["inv_depths"][1].shape[-1] * 2 == out["inv_depths"][0].shape[-1]      # maps stay at their own resolution (a8 fused)
            for j, pz in enumerate(out["poses"]):
                # PoseNet runs on the library's fp32 convolutions (models.require_fp32_library_convolutions)
                assert torch.allclose(pz.mat.detach().cpu(), z["pose%d_step0" % j], atol=2e-6, rtol=1e-4)
            worst = ("", 0.0)
            for prefix, net in (("depth.", model.depth_net), ("pose.", model.pose_net)):
                for k, p in net.named_parameters():
                    norm_ref = float(z["g0norm/" + prefix + k])
                    flat = p.grad.reshape(-1)
                    norm_got = float(flat.double().norm())
                    # absolute floor: a Conv3d bias in front of Conv2d + GroupNorm has a gradient that is zero up to border effects
                    # -- what is left is mostly rounding (7.7e-4 for pack3.conv3d.bias against ~1 for its neighbours)
                    tol = 5e-3 * norm_ref + 5e-6
                    worst = max(worst, (prefix + k, abs(norm_got - norm_ref) / tol), key=lambda t: t[1])
                    assert abs(norm_got - norm_ref) <= tol, (prefix + k, norm_got, norm_ref)
            print("worst gradient-norm error/bound after step 0: %s %.3f" % worst)
        opt.step()
        # Adam's first steps move every element by about lr * sign(g): elements whose gradient is within rounding of zero may
        # go the other way, so the bar is a fraction of samples within 10 % of lr, not a norm
        named = {("depth." + k): p for k, p in model.depth_net.named_parameters()}
        named.update({("pose." + k): p for k, p in model.pose_net.named_parameters()})
        keys = [k for k in z if k.startswith("p%d/" % step)]
        assert len(keys) >= 8
        for key in keys:
            p = named[key.split("/", 1)[1]].detach().reshape(-1)
            n = z[key].numel()
            got = p[strided_index(p.numel(), n).to(DEV)].cpu()
            close = ((got - z[key]).abs() <= 2e-5).float().mean().item()
            assert close >= 0.97, (key, close)
