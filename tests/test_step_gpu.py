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
            # step 1 sits one Adam update apart: the first update moves EVERY element by +-lr, also those whose gradient is
            # rounding noise, so the small smoothness term (0.05 % of the loss) differs by ~1 % between any two runs (measured
            # on the B200: 7.85e-5 .. 7.92e-5 against the reference's 7.84e-5) while the loss stays within 1e-3
            tol = 1e-3 if (step == 0 or key == "photometric_loss") else 3e-2
            assert abs(got - ref) <= tol * abs(ref), (key, got, ref)
        if step == 0:
            assert ((out["inv_depths"][0].detach().cpu() - z["inv_depth0_step0"]).abs() / z["inv_depth0_step0"].abs()).max() < 1e-3
            assert out["inv_depths"][1].shape[-1] * 2 == out["inv_depths"][0].shape[-1]      # maps stay at their own resolution (a8 fused)
            for j, pz in enumerate(out["poses"]):
                # PoseNet runs on the library's fp32 convolutions (models.require_fp32_library_convolutions)
                assert torch.allclose(pz.mat.detach().cpu(), z["pose%d_step0" % j], atol=2e-6, rtol=1e-4)
            # Gradient norms against the reference's.  The bound is NOT a rounding bound: a handful of pixels whose two best
            # candidates of the per-pixel minimum tie to ~1e-6 resolve differently from run to run (split-K atomics reorder the
            # forward sums) and on a 64x96 image each of them moves a decoder gradient by ~1e-3 of its norm; the tight
            # comparison of the two weight storages below (same run-to-run noise, no reference) is the rounding-level check.
            devs = []
            for prefix, net in (("depth.", model.depth_net), ("pose.", model.pose_net)):
                for k, p in net.named_parameters():
                    norm_ref = float(z["g0norm/" + prefix + k])
                    norm_got = float(p.grad.reshape(-1).double().norm())
                    devs.append((abs(norm_got - norm_ref) / (1.5e-2 * norm_ref + 5e-6), prefix + k, norm_got, norm_ref))
            devs.sort(reverse=True)
            print("largest gradient-norm deviations from the reference after step 0 (error / bound, name, ours, reference):")
            for d in devs[:6]:
                print("   %.3f %s %.6e %.6e" % d)
            assert devs[0][0] <= 1.0, devs[:3]
        opt.step()
        # Adam's first steps move every element by about lr * sign(g): elements whose gradient is within rounding of zero may
        # go the other way, so the bar is a fraction of samples within 10 % of lr, not a norm
        named = {("depth." + k): p for k, p in model.depth_net.named_parameters()}
        named.update({("pose." + k): p for k, p in model.pose_net.named_parameters()})
        keys = [k for k in z if k.startswith("p%d/" % step)]
        assert len(keys) >= 8
        worst, pooled = [], [0, 0]
        for key in keys:
            if float(z["g0norm/" + key.split("/", 1)[1]]) < 1e-6:
                continue        # gradient = rounding noise (a bias in front of a one-channel-per-group GroupNorm): Adam moves it +-lr at random
            p = named[key.split("/", 1)[1]].detach().reshape(-1)
            n = z[key].numel()
            got = p[strided_index(p.numel(), n).to(DEV)].cpu()
            close = ((got - z[key]).abs() <= 2e-5).float().mean().item()
            worst.append((close, key))
            pooled[0] += int(((got - z[key]).abs() <= 2e-5).sum())
            pooled[1] += n
            # step 1 is a second update on parameters that already differ by +-lr wherever step 0's gradient was noise, and its
            # own gradient went through the flipped network: per tensor the fraction falls to 0.89-0.99 on the first layers and
            # to 14/16 on a 16-element PoseNet bias (measured on the B200 over six runs of the three optimizer variants; the
            # step-1 loss itself moves by up to 7e-4 relative between runs); the pooled fraction was 0.943 .. 0.99, so step 1 is
            # a sanity bound (a wrong optimizer moves EVERY element) and step 0 is the rounding-level check
            assert close >= (0.97 if step == 0 else (0.75 if n >= 64 else 0.6)), (key, close)
        assert pooled[0] >= (0.97 if step == 0 else 0.90) * pooled[1], pooled
        worst.sort()
        print("step %d: smallest fractions of parameter samples within 2e-5 of the reference:" % step, [(round(c, 4), k) for c, k in worst[:3]])


def test_stored_and_plain_weight_storage_give_the_same_gradients_and_step():
    """PackNet01 twice on the kernels of this repo with a FIXED output gradient: convolution weights stored in the engine's layout
    (tiles from the optimizer launch, weight gradients accumulated in the flat buffer) against plain OIHW storage (packed per
    call, gradients unpacked).  Same arithmetic: every gradient agrees to the run-to-run noise of the split-K / weight-gradient
    atomics (measured with tools/determinism_probe.py: median 3e-5, worst 6e-5), and so do the parameters after one optimizer
    step wherever |g| is above that noise.  (Through the photometric loss the comparison would be meaningless at this level:
    its sign() terms turn the 5e-6 run-to-run noise of the depth maps into ~1 % of every gradient, in the reference as here.)"""
    from packnet_sfm_b200 import optim, synthetic
    from packnet_sfm_b200.networks import native_conv_weights
    x = synthetic.make_frames(2, 64, 96, seed=9)["rgb"].to(DEV)
    g = torch.Generator().manual_seed(4)
    gys = None
    res = []
    for stored in (True, False):
        net = _model().depth_net
        opt = optim.FlatAdam(net.parameters(), lr=2e-4, native=native_conv_weights(net, (64, 96)) if stored else ())
        assert (len(opt.natives) > 40) == stored
        opt.zero_grad()
        outs = net(x)["inv_depths"]
        if gys is None:
            gys = [(torch.rand(o.shape, generator=g) - 0.5).to(DEV) for o in outs]
        torch.autograd.backward(outs, gys)
        opt.collect_grads()
        torch.cuda.synchronize()
        grads = {k: p.grad.detach().clone() for k, p in net.named_parameters()}
        before = {k: p.detach().clone() for k, p in net.named_parameters()}
        opt.step()
        res.append(([o.detach().clone() for o in outs], grads, {k: p.detach() - before[k] for k, p in net.named_parameters()}))
    (o1, g1, d1), (o0, g0, d0) = res
    for a, b in zip(o1, o0):
        assert rel_l2(a, b) < 5e-5
    devs = sorted(((float((g1[k].double() - g0[k].double()).norm() / (g0[k].double().norm() + 1e-12)), k) for k in g0
                   if float(g0[k].double().norm()) > 1e-6), reverse=True)
    print("largest relative gradient differences stored vs plain:", [(round(d, 7), k) for d, k in devs[:5]])
    assert devs[0][0] < 5e-4, devs[:5]
    assert sorted(d for d, _ in devs)[len(devs) // 2] < 1e-4
    for k in d0:
        big = g0[k].abs() > 1e-2 * g0[k].abs().max()
        if big.any():      # where the gradient is above the noise both runs step the same way (first Adam step: -lr * sign(g))
            assert float(((d1[k] - d0[k]).abs()[big] > 1e-5).float().mean()) < 0.01, k
