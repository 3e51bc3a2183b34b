"""Data-parallel plumbing of the hot path: frames shard over ranks, ONE flat-buffer all-reduce on the
gradients per step (replaces DistributedDataParallel's bucketed reducer, engine/train.py:163-164).
torch.distributed (NCCL over NVLink/NVSwitch on the GPU box, gloo in CPU tests) is plumbing only."""
import torch
import torch.distributed as dist


class FlatGradAllReduce:
    """Owns one contiguous fp32 buffer; every trainable parameter's .grad is a view into it, so backward writes
    gradients in place and a single collective reduces all of them."""

    def __init__(self, params, dtype=torch.float32):
        self.params = [p for p in params if p.requires_grad]
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, dtype=dtype, device=dev)
        off = 0
        for p in self.params:
            p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def zero(self):
        self.flat.zero_()
        off = 0
        for p in self.params:  # re-attach views in case an optimizer set grads to None
            if p.grad is None or p.grad.data_ptr() != self.flat[off:].data_ptr():
                p.grad = self.flat[off:off + p.numel()].view_as(p)
            off += p.numel()

    def allreduce(self, average=True):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
            if average:
                self.flat.div_(dist.get_world_size())
        return self.flat


def shard_frames(n_frames_global, rank, world):
    """Contiguous per-rank frame range (frames of one clip never split: shard by clip at the caller)."""
    per = n_frames_global // world
    return rank * per, (rank + 1) * per
