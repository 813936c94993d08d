"""Worker of tests/test_ddp_gpu.py (launched with torch.distributed.run, one rank per GPU, NCCL): the rank-averaged gradients
of the data-parallel step -- bucket all-reduces launched from the backward, optim.FlatAdam -- against the single-process
gradient of the concatenated batch (Horovod average, trainers/horovod_trainer.py:46-48), then one optimizer step
and identical replicas.  Not collected by pytest."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def build(seed, H, W, buckets):
    from packnet_sfm_b200 import optim
    from packnet_sfm_b200.models import SelfSupModel
    from packnet_sfm_b200.networks import gradient_buckets, native_conv_weights
    torch.manual_seed(seed)
    model = SelfSupModel(flip_lr_prob=0.0).cuda().train()
    opt = optim.FlatAdam([{"params": list(model.depth_net.parameters())}, {"params": list(model.pose_net.parameters())}], lr=2e-4,
                         native=native_conv_weights(model.depth_net, (H, W)),
                         buckets=gradient_buckets(model.depth_net, (H, W)) if buckets else None)
    return model, opt


def batch_of(fr, sl):
    b = {"rgb": fr["rgb"][sl].cuda(), "rgb_context": [c[sl].cuda() for c in fr["rgb_context"]], "intrinsics": fr["intrinsics"][sl].cuda()}
    b["rgb_original"], b["rgb_context_original"] = b["rgb"], b["rgb_context"]
    return b


def objective(model, fr, sl, gys):
    """A smooth stand-in for the loss: mean over the samples of <disp_i, gy_i> + <pose vectors, 1>.  Linear in the per-sample
    terms like the reference's loss mean, but without the photometric loss's sign() terms, which turn the 5e-6
    run-to-run noise of the depth maps into ~1 % of every gradient (tools/determinism_probe.py) and would drown the comparison."""
    rgb, ctx = fr["rgb"][sl].cuda(), [c[sl].cuda() for c in fr["rgb_context"]]
    n = rgb.shape[0]
    outs = model.depth_net(rgb)["inv_depths"]
    obj = sum((o * g[sl].cuda()).sum() for o, g in zip(outs, gys)) / n
    return obj + model.pose_net(rgb, ctx).sum() / n


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    torch.cuda.set_stream(torch.cuda.Stream())
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from packnet_sfm_b200 import parallel, synthetic
    H, W, per = 64, 96, 2
    fr = synthetic.make_frames(per * world, H, W, seed=17)
    g = torch.Generator().manual_seed(23)
    gys = [torch.rand(per * world, 1, H >> i, W >> i, generator=g) - 0.5 for i in range(4)]
    model, opt = build(5, H, W, buckets=True)
    parallel.broadcast_parameters(model)
    opt.repack()                                   # the broadcast wrote the parameters through PyTorch
    assert len(opt._bucket_ranges) == 2
    opt.zero_grad()
    objective(model, fr, parallel.shard_batch(per * world, rank, world), gys).backward()
    launched_in_backward = list(opt._launched)
    opt.allreduce_mean()
    torch.cuda.synchronize()
    grads = {k: p.grad.detach().clone() for k, p in model.named_parameters()}
    opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in model.parameters()])
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    ok = all(torch.equal(gathered[0], t) for t in gathered)
    worst, med = ("", 0.0), 0.0
    if rank == 0:
        ref_model, ref_opt = build(5, H, W, buckets=False)
        ref_opt.zero_grad()
        objective(ref_model, fr, slice(0, per * world), gys).backward()
        ref_opt.collect_grads()
        torch.cuda.synchronize()
        errs = []
        for k, p in ref_model.named_parameters():
            a, b = grads[k].double(), p.grad.detach().double()
            if float(b.norm()) < 1e-6:
                continue                           # pure rounding noise (a bias in front of a one-channel-per-group GroupNorm)
            errs.append((float((a - b).norm() / b.norm()), k))
        errs.sort(reverse=True)
        worst, med = (errs[0][1], errs[0][0]), errs[len(errs) // 2][0]
        print("DDP launched_in_backward=%s replicas_identical=%s worst_grad_rel_l2=%.3e (%s) median=%.3e" % (
            launched_in_backward, ok, worst[1], worst[0], med), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    # fp32 sums in another order (two shard means vs one batch mean, atomics): the network-alone noise is ~5e-5
    sys.exit(0 if (ok and all(launched_in_backward) and worst[1] < 1e-3 and med < 2e-4) else 1)


if __name__ == "__main__":
    main()
