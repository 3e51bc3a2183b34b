"""world_size-2 gloo test of the flat gradient all-reduce used by the N>1 path."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _worker(rank, world, port):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from maggie_b200.dp import FlatGradAllReduce, shard_frames

    torch.manual_seed(0)
    lin = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    fr = FlatGradAllReduce(lin.parameters())
    x = torch.arange(16.0).view(4, 4)
    lo, hi = shard_frames(4, rank, world)
    fr.zero()
    lin(x[lo:hi]).sum().backward()
    fr.allreduce(average=False)
    got = fr.flat.clone()
    # single-process reference over the full batch
    ref = torch.nn.Sequential(torch.nn.Linear(4, 3), torch.nn.Linear(3, 2))
    ref.load_state_dict(lin.state_dict())
    ref(x).sum().backward()
    want = torch.cat([p.grad.flatten() for p in ref.parameters()])
    assert torch.allclose(got, want, atol=1e-5), (got, want)
    assert all(p.grad.data_ptr() >= fr.flat.data_ptr() for p in lin.parameters())
    _sync_bn_switches(rank, world)
    dist.destroy_process_group()


def _sync_bn_switches(rank, world):
    """Host logic of the SyncBatchNorm-equivalent statistics exchange: which containers exchange, and over what."""
    from maggie_b200 import dense, dp
    plain, conv = torch.nn.BatchNorm2d(8), torch.nn.SyncBatchNorm(8)
    assert dense.sync_group(plain) is None
    assert dense.sync_group(conv) is dist.group.WORLD
    model = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 1), plain)
    assert not dense.sync_bn_active(model)
    assert dense.sync_bn_active(torch.nn.SyncBatchNorm.convert_sync_batchnorm(model))
    dp.set_sync_bn(True)
    assert dense.sync_group(plain) is dist.group.WORLD and dense.sync_bn_active(torch.nn.Sequential(plain))
    dp.set_sync_bn(False)
    assert dense.sync_group(plain) is None
    t = torch.full((5,), float(rank + 1))
    assert torch.equal(dense.exchange(t, dist.group.WORLD), torch.full((5,), 3.0))


def test_flat_allreduce_matches_full_batch_gradient():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port), nprocs=2, join=True)
