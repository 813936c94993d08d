"""Optimizer + gradient bucket of the training step on flat buffers (include/packnet_b200.h: pn_adam_step).

Replaces, for the hot path, torch.optim.Adam over the reference's 'Depth' / 'Pose' parameter groups
(packnet_sfm/models/model_wrapper.py:128-166), Horovod's DistributedOptimizer all-reduce of the gradients
(trainers/horovod_trainer.py:46-48,92-93) and the per-step weight re-layouts of the convolution engine:

  * every parameter is a VIEW into one flat fp32 buffer (so are its gradient and both Adam moments); one kernel launch
    updates all 129.9 M elements;
  * a convolution weight that feeds the tensor-core engine (`native`) is stored as [Cout][tap][kpad] -- the parameter keeps
    its [Cout,Cin,k,k] shape, names and values (state_dict / load_state_dict / checkpoints are unchanged), only its strides
    differ -- which is the layout the weight-gradient kernel accumulates: its gradient is written straight into the flat
    gradient buffer, and the same launch that updates it writes the bf16 hi/lo forward tiles the next step's forward AND
    data-gradient convolutions read (functional._Conv2d);
  * the flat gradient buffer is what the data-parallel all-reduce averages (one NCCL call, or a few reverse-order
    buckets overlapped with the backward: `allreduce_mean` / `grad_ready_hook`).

Not a torch.optim.Optimizer subclass on purpose: there is no per-parameter state dict to keep consistent, and the step must
be capturable in a CUDA graph (step count, bias corrections and learning rates live in device memory)."""
import ctypes

import numpy as np
import torch
import torch.distributed as dist

from . import _lib

BLOCK = 2048          # PN_ADAM_BLOCK
MAX_GROUPS = 4


class ConvSeg(ctypes.Structure):      # pn_adam_conv_seg
    _fields_ = [("offset", ctypes.c_int64), ("packed_offset", ctypes.c_int64),
                ("cout", ctypes.c_int32), ("taps", ctypes.c_int32), ("kpad", ctypes.c_int32), ("rows_pad", ctypes.c_int32)]


def _round_up(x, a):
    return (x + a - 1) // a * a


class NativeWeight:
    """What functional._Conv2d needs to use a stored convolution weight without touching its fp32 values."""

    def __init__(self, owner, index, cout, cin, ksize, kpad, rows_pad, hi, lo, grad_flat):
        self.owner, self.index = owner, index
        self.cout, self.cin, self.ksize, self.kpad, self.rows_pad = cout, cin, ksize, kpad, rows_pad
        self.hi, self.lo = hi, lo                  # bf16 forward tiles [kpad/64][taps][rows_pad][64]
        self.grad_flat = grad_flat                 # fp32 [cout*taps*kpad]: slice of the flat gradient buffer
        self.version = -1                          # parameter._version the tiles were written at

    def accepts(self, cin_tensor):
        """the activation may carry the channel padding of the operand (multiple of 8) but must need the same 64-chunks"""
        return self.cin <= cin_tensor <= self.kpad and (cin_tensor + 63) // 64 == self.kpad // 64


class FlatAdam:
    """Adam on flat buffers.  `params`: an iterable of parameters or of torch-style groups ({'params': [...], 'lr': ...,
    'weight_decay': ...}); `native`: the 4-D convolution weights to store in the engine's layout (networks.native_conv_weights);
    `buckets`: lists of native weights in the order their gradients complete during the backward (networks.gradient_buckets):
    each list becomes one contiguous range at the front of the flat buffers whose all-reduce is launched from the backward
    as soon as its last weight gradient has been enqueued; everything else is reduced by allreduce_mean() after the backward."""

    def __init__(self, params, lr=2e-4, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, native=(), buckets=None):
        groups = list(params)
        if groups and not isinstance(groups[0], dict):
            groups = [{"params": groups}]
        assert 0 < len(groups) <= MAX_GROUPS, "1..%d parameter groups" % MAX_GROUPS
        self.param_groups = []
        for g in groups:
            g = dict(g)
            g["params"] = [p for p in g["params"] if p.requires_grad]
            g.setdefault("lr", lr)
            g.setdefault("weight_decay", weight_decay)
            self.param_groups.append(g)
        self.betas, self.eps = betas, eps
        native_ids = {id(p) for p in native}
        every = [p for g in self.param_groups for p in g["params"]]
        assert every, "no trainable parameters"
        assert len({id(p) for p in every}) == len(every), "a parameter appears twice"
        dev = every[0].device
        assert all(p.device == dev and p.dtype == torch.float32 for p in every), "fp32 parameters on one device"
        lib = _lib.lib()
        # ---- layout: [bucket 0 natives | bucket 1 natives | ... | other natives | group 0 plain | group 1 plain | ...],
        # every stored weight and every group padded to whole blocks
        group_of = {id(p): gi for gi, g in enumerate(self.param_groups) for p in g["params"]}
        ordered, bucket_sizes = [], []
        for bl in (buckets or []):
            bl = [p for p in bl if id(p) in native_ids and id(p) not in {id(q) for q in ordered}]
            if bl:
                ordered += bl
                bucket_sizes.append(len(bl))
        ordered += [p for p in every if id(p) in native_ids and id(p) not in {id(q) for q in ordered}]
        self._slots = []                       # (param, offset, numel_storage, group, native index or -1)
        block_info = []
        segs, packed_bytes = [], 0
        off = 0
        self._bucket_ranges, self._bucket_of = [], {}
        bstart, bcount, bi = 0, 0, 0
        for p in ordered:
            assert p.dim() == 4 and p.shape[2] == p.shape[3], "native weights are [Cout,Cin,k,k]"
            gi = group_of[id(p)]
            cout, cin, k, _ = p.shape
            kpad, rows_pad = _round_up(cin, 64), int(lib.pn_conv2d_rows_pad(cout))
            n = cout * k * k * kpad
            seg = ConvSeg(off, packed_bytes, cout, k * k, kpad, rows_pad)
            if bi < len(bucket_sizes):
                self._bucket_of[len(segs)] = bi
            self._slots.append((p, off, n, gi, len(segs)))
            segs.append(seg)
            nb = _round_up(n, BLOCK) // BLOCK
            block_info += [(gi << 16) | len(segs)] * nb
            off += nb * BLOCK
            packed_bytes += _round_up((kpad // 64) * k * k * rows_pad * 128, 1024)
            bcount += 1
            while bi < len(bucket_sizes) and bcount == bucket_sizes[bi]:
                self._bucket_ranges.append((bstart, off, bucket_sizes[bi]))
                bstart, bcount, bi = off, 0, bi + 1
        self._tail_start = self._bucket_ranges[-1][1] if self._bucket_ranges else 0
        for gi, g in enumerate(self.param_groups):
            for p in g["params"]:
                if id(p) in native_ids:
                    continue
                self._slots.append((p, off, p.numel(), gi, -1))
                off += _round_up(p.numel(), 4)                      # 16-byte aligned views
            end = _round_up(off, BLOCK)
            block_info += [gi << 16] * ((end - len(block_info) * BLOCK) // BLOCK)
            off = end
        assert len(segs) < 0xFFFF
        self.numel = off
        assert len(block_info) * BLOCK == off
        self.flat_param = torch.zeros(off, dtype=torch.float32, device=dev)
        self.flat_grad = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros(off, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(off, dtype=torch.float32, device=dev)
        self.packed_hi = torch.zeros(max(packed_bytes, 16), dtype=torch.uint8, device=dev)     # tile rows >= Cout stay zero
        self.packed_lo = torch.zeros(max(packed_bytes, 16), dtype=torch.uint8, device=dev)
        self._block_info = torch.from_numpy(np.asarray(block_info, dtype=np.int32)).to(dev)
        seg_bytes = b"".join(bytes(s) for s in segs) or bytes(ctypes.sizeof(ConvSeg))
        self._segs = torch.frombuffer(bytearray(seg_bytes), dtype=torch.uint8).to(dev)
        hyper = np.zeros(16, dtype=np.float32)
        hyper[1], hyper[2], hyper[3] = betas[0], betas[1], eps
        self._hyper_host = hyper
        self.hyper = torch.from_numpy(hyper.copy()).to(dev)
        self._write_group_hyper()
        # ---- re-point the parameters (values preserved) and their gradients
        self.natives = []
        self._plain = []                        # (param, grad view)
        with torch.no_grad():
            for p, o, n, gi, si in self._slots:
                if si < 0:
                    view = self.flat_param[o:o + n].view(p.shape)
                    view.copy_(p.data)
                    p.data = view
                    self._plain.append((p, self.flat_grad[o:o + n].view(p.shape)))
                else:
                    cout, cin, k, _ = p.shape
                    s = segs[si]
                    strides = (s.taps * s.kpad, 1, k * s.kpad, s.kpad)
                    view = torch.as_strided(self.flat_param, (cout, cin, k, k), strides, o)
                    view.copy_(p.data)
                    p.data = view
                    p.grad = torch.as_strided(self.flat_grad, (cout, cin, k, k), strides, o)
                    tile_elems = (s.kpad // 64) * s.taps * s.rows_pad * 64
                    hi = self.packed_hi[s.packed_offset:s.packed_offset + 2 * tile_elems].view(torch.bfloat16)
                    lo = self.packed_lo[s.packed_offset:s.packed_offset + 2 * tile_elems].view(torch.bfloat16)
                    nat = NativeWeight(self, si, cout, cin, k, s.kpad, s.rows_pad, hi, lo, self.flat_grad[o:o + n])
                    p._pn_native = nat
                    self.natives.append((p, nat))
        self._collected = False
        self._pending = [n for (_, _, n) in self._bucket_ranges]     # weight gradients still missing per bucket, this step
        self._works = []
        self._launched = [False] * len(self._bucket_ranges)
        self.repack()

    # ------------------------------------------------------------------------------------------------------------
    def _write_group_hyper(self):
        h = self.hyper.cpu().numpy()
        for gi, g in enumerate(self.param_groups):
            h[8 + 2 * gi], h[9 + 2 * gi] = g["lr"], g["weight_decay"]
        self.hyper.copy_(torch.from_numpy(h))

    def set_lr(self, lrs):
        """StepLR and friends (model_wrapper.py:150-166): new learning rate(s), one value or one per group.  A device copy,
        outside any captured graph -- the graph reads the values at replay time."""
        lrs = [lrs] * len(self.param_groups) if not isinstance(lrs, (list, tuple)) else lrs
        for g, v in zip(self.param_groups, lrs):
            g["lr"] = float(v)
        self._write_group_hyper()

    def _call(self, update):
        lib = _lib.lib()
        lib.pn_adam_step.argtypes = [ctypes.c_void_p] * 4 + [ctypes.c_int64] + [ctypes.c_void_p] * 5 + [ctypes.c_int, ctypes.c_void_p]
        lib.pn_adam_step.restype = ctypes.c_int
        _lib.check(lib.pn_adam_step(_lib.ptr(self.flat_param), _lib.ptr(self.flat_grad), _lib.ptr(self.exp_avg),
                                    _lib.ptr(self.exp_avg_sq), self.numel, _lib.ptr(self._block_info), _lib.ptr(self._segs),
                                    _lib.ptr(self.hyper), _lib.ptr(self.packed_hi), _lib.ptr(self.packed_lo), int(update),
                                    _lib.current_stream()), "pn_adam_step")

    def repack(self):
        """(Re)write the forward tiles of every stored weight from the fp32 values (construction, load_state_dict, manual
        re-initialisation: anything that wrote the parameters through PyTorch)."""
        if self.natives:
            _lib.require_f32(self.flat_param)
            self._call(False)
        for p, nat in self.natives:
            nat.version = p._version

    # ------------------------------------------------------------------------------------------------------------
    def zero_grad(self, set_to_none=True):
        """Plain parameters: grad = None (autograd assigns a fresh tensor, collected into the bucket by `collect_grads`).
        Stored convolution weights keep their bucket view: the weight-gradient kernel overwrites it every step."""
        for p, _ in self._plain:
            p.grad = None
        self._collected = False
        self._pending = [n for (_, _, n) in self._bucket_ranges]
        self._launched = [False] * len(self._bucket_ranges)

    def collect_grads(self):
        """Gather the gradients autograd produced for the plain parameters into the flat buffer (one multi-tensor copy)."""
        if self._collected:
            return
        src, dst = [], []
        for p, v in self._plain:
            if p.grad is None:
                v.zero_()
            elif p.grad.data_ptr() != v.data_ptr():
                src.append(p.grad)
                dst.append(v)
        if src:
            torch._foreach_copy_(dst, src)
        for p, v in self._plain:
            p.grad = v
        self._collected = True

    @staticmethod
    def _distributed(group=None):
        return dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1

    def _reduce_range(self, start, end, group=None):
        """launch the averaging all-reduce of flat_grad[start:end]; NCCL orders it after the work enqueued so far on the
        current stream and runs it on its own stream, i.e. concurrently with whatever the backward enqueues next"""
        if end <= start:
            return
        view = self.flat_grad[start:end]
        if dist.get_backend(group) == "nccl":
            self._works.append((dist.all_reduce(view, op=dist.ReduceOp.AVG, group=group, async_op=True), None))
        else:
            self._works.append((dist.all_reduce(view, op=dist.ReduceOp.SUM, group=group, async_op=True), view))

    def grad_ready(self, nat, group=None):
        """Called by functional._Conv2d.backward after it enqueued the weight gradient of a stored weight: when that
        completes a bucket, the bucket's all-reduce starts NOW and overlaps the rest of the backward (the reference's Horovod
        optimizer does the same per tensor, trainers/horovod_trainer.py:46-48)."""
        bi = self._bucket_of.get(nat.index)
        if bi is None or not self._distributed(group):
            return
        self._pending[bi] -= 1
        if self._pending[bi] == 0 and not self._launched[bi]:
            self._launched[bi] = True
            self._reduce_range(self._bucket_ranges[bi][0], self._bucket_ranges[bi][1], group)

    def allreduce_mean(self, group=None):
        """Average the gradients over the ranks (Horovod's op=Average): the buckets not yet launched from the backward plus
        the tail of the flat buffer (remaining stored weights and all plain parameters), then wait for every collective."""
        if not self._distributed(group):
            return
        self.collect_grads()
        for bi, (start, end, _) in enumerate(self._bucket_ranges):
            if not self._launched[bi]:
                self._launched[bi] = True
                self._reduce_range(start, end, group)
        self._reduce_range(self._tail_start, self.numel, group)
        world = dist.get_world_size(group)
        for work, view in self._works:
            work.wait()
            if view is not None:
                view.div_(world)
        self._works = []

    def step(self):
        self.collect_grads()
        _lib.require_f32(self.flat_param, self.flat_grad)
        self._call(True)
        self._collected = False

    def nbytes(self):
        return self.numel * 4

    # ------------------------------------------------------------------------------------------------------------
    def state_dict(self):
        return {"hyper": self.hyper.cpu(), "exp_avg": self.exp_avg.cpu(), "exp_avg_sq": self.exp_avg_sq.cpu(),
                "param_groups": [{k: v for k, v in g.items() if k != "params"} for g in self.param_groups]}

    def load_state_dict(self, sd):
        self.hyper.copy_(sd["hyper"])
        self.exp_avg.copy_(sd["exp_avg"])
        self.exp_avg_sq.copy_(sd["exp_avg_sq"])
        for g, s in zip(self.param_groups, sd["param_groups"]):
            g.update(s)
        self._write_group_hyper()
