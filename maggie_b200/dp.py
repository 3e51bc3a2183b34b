"""Data-parallel plumbing of the hot path: frames shard over ranks, ONE flat-buffer all-reduce on the
gradients per step (replaces DistributedDataParallel's bucketed reducer, engine/train.py:163-164).
torch.distributed (NCCL over NVLink/NVSwitch on the GPU box, gloo in CPU tests) is plumbing only."""
import torch
import torch.distributed as dist


class FlatGradAllReduce:
    """One flat fp32 buffer for the gradients of all trainable parameters and ONE collective per step.

    `zero()` drops the gradients (`p.grad = None`): autograd then hands each parameter its freshly computed gradient
    tensor instead of launching an accumulate kernel per parameter (~300 tiny launches per step).  `allreduce()` packs
    the gradients into the flat buffer with multi-tensor copies, reduces it once (averaging inside the collective), and
    leaves every `p.grad` as a view into the buffer.  With one rank nothing is packed at all.

    `bank` (the model's `WeightBank`, `model.bank`): the grouped weight-preparation backward (K0) then writes the gradients
    of all its conv weights - 95 % of the bytes - straight into the head of the flat buffer, in the bank's own order, so
    that nothing is left to pack for them (`offsets[i]` is the position of parameter i; the parameter ORDER, which numbers
    the optimizer state as torch.optim does, is unchanged)."""

    def __init__(self, params, dtype=torch.float32, bank=None):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=dtype, device=dev)
        self.no_grad = ()
        head = {}
        if bank is not None and bank.entries and dtype == torch.float32:
            mine, off = {id(p) for p in self.params}, 0
            for e in bank.entries:
                head[id(e.w)] = off
                off += e.w.numel()
            if all(k in mine for k in head) and len(head) == len(bank.entries):
                bank.grad_target = self.flat[:off]
            else:
                head = {}
        self.offsets, off = [], sum(p.numel() for p in self.params if id(p) in head)
        for p in self.params:
            if id(p) in head:
                self.offsets.append(head[id(p)])
            else:
                self.offsets.append(off)
                off += p.numel()
        self.views = [self.flat[o:o + p.numel()].view_as(p) for o, p in zip(self.offsets, self.params)]

    def zero(self):
        for p in self.params:
            p.grad = None

    def pack(self):
        """Gradients -> flat buffer (parameters without a gradient contribute zeros); p.grad becomes the view."""
        # (indices of the parameters that received no gradient this step: the optimizer leaves them alone, as torch does)
        self.no_grad = tuple(i for i, p in enumerate(self.params) if p.grad is None)
        have = [(v, p.grad) for v, p in zip(self.views, self.params) if p.grad is not None and p.grad.data_ptr() != v.data_ptr()]
        if self.no_grad:
            for v, p in zip(self.views, self.params):
                if p.grad is None:
                    v.zero_()
        if have:
            torch._foreach_copy_([v for v, _ in have], [g for _, g in have])
        for v, p in zip(self.views, self.params):
            if p.grad is not None:
                p.grad = v
        return self.flat

    def allreduce(self, average=True):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            self.pack()
            if average and dist.get_backend() == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG)      # 1 / world inside the collective: no extra pass
            else:
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
                if average:
                    self.flat.div_(dist.get_world_size())
        return self.flat


def shard_frames(n_frames_global, rank, world):
    """Contiguous per-rank frame range (frames of one clip never split: shard by clip at the caller)."""
    per = n_frames_global // world
    return rank * per, (rank + 1) * per


def set_sync_bn(group=True):
    """SyncBatchNorm-equivalent training (`model.sync_bn: true`, engine/train.py:160-161): every BatchNorm takes its
    batch statistics over the frames / active sites of all ranks of `group`.  Equivalent to running the model after
    `nn.SyncBatchNorm.convert_sync_batchnorm(model)`, which the native path also honours.  See maggie_b200/dense.py."""
    from . import dense
    dense.set_sync_bn(group)
