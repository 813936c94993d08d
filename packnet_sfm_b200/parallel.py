"""Data parallelism for the self-supervised step: one process per GPU, pure batch split, ONE all-reduce of a
flat fp32 gradient bucket per step (replaces Horovod's per-tensor DistributedOptimizer all-reduces,
packnet_sfm/trainers/horovod_trainer.py:46-48,92-93) plus an explicit parameter broadcast at start (the
reference relies on identical seeds, model_wrapper.py:44,374-379).

torch.distributed is the plumbing (backend nccl over NVLink 5 / NVSwitch on the GPU box, gloo in the CPU
tests).  The path has no other exchange step: samples are independent."""
import torch
import torch.distributed as dist


class FlatBucket:
    """One contiguous gradient buffer for the single all-reduce of a step.

    Autograd ASSIGNS fresh gradient tensors (every p.grad is None when the backward starts), `pack()` gathers them into
    the flat buffer with one multi-tensor copy and re-points every p.grad at its slice, so the all-reduce and the
    optimizer work on the bucket.  (Keeping p.grad pinned to bucket views instead makes autograd launch one `add_` per
    parameter -- 216 launches per step for PackNet01 -- plus a 520 MB memset.)  With a single rank nothing is packed."""

    def __init__(self, params):
        self.params = [p for p in params if p.requires_grad]
        assert self.params, "no trainable parameters"
        dev, dt = self.params[0].device, self.params[0].dtype
        self.numel = sum(p.numel() for p in self.params)
        self.flat_grad = torch.zeros(self.numel, dtype=dt, device=dev)
        self.views = []
        off = 0
        for p in self.params:
            n = p.numel()
            self.views.append(self.flat_grad[off:off + n].view_as(p))
            off += n

    def zero_grad(self):
        for p in self.params:
            p.grad = None

    def pack(self):
        """Gather the gradients autograd produced into the bucket; afterwards every p.grad IS its bucket slice."""
        src, dst = [], []
        for p, v in zip(self.params, self.views):
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in zip(self.params, self.views):
            p.grad = v

    def nbytes(self):
        return self.numel * self.flat_grad.element_size()

    def allreduce_mean(self, group=None, async_op=False):
        """Average the gradients over the ranks (Horovod's default op=Average)."""
        if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
            return None
        self.pack()
        world = dist.get_world_size(group)
        if dist.get_backend(group) == "nccl":
            return dist.all_reduce(self.flat_grad, op=dist.ReduceOp.AVG, group=group, async_op=async_op)
        work = dist.all_reduce(self.flat_grad, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        if async_op:
            work.wait()
        self.flat_grad.div_(world)
        return None


def broadcast_parameters(module, src=0, group=None):
    """One flat broadcast of all parameters and buffers from rank `src`."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    tensors = [p.data for p in module.parameters()] + [b.data for b in module.buffers()]
    if not tensors:
        return
    flat = torch.cat([t.reshape(-1).float() for t in tensors])
    dist.broadcast(flat, src=src, group=group)
    off = 0
    for t in tensors:
        n = t.numel()
        t.copy_(flat[off:off + n].view_as(t).to(t.dtype))
        off += n


def shard_batch(batch_size, rank, world_size):
    """Rank r takes samples [r*B_local, (r+1)*B_local) (DistributedSampler, model_wrapper.py:569-573)."""
    assert batch_size % world_size == 0, "global batch must divide the number of ranks"
    per = batch_size // world_size
    return slice(rank * per, (rank + 1) * per)
