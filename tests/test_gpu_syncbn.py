"""SyncBatchNorm-equivalent statistics exchange (engine/train.py:160-161): two ranks, each with one shard of
the batch and `nn.SyncBatchNorm` containers, must reproduce ONE process that runs the whole batch with local statistics -
outputs, data gradients, summed parameter gradients and running statistics.  Both ranks share cuda:0 (NCCL refuses two
ranks on one device, so the host-side group is gloo); the statistics travel through the K15 peer-memory exchange windows
(CUDA IPC mappings of each other's window - the same kernel and protocol as across NVLink) and, in the second test,
through the collective fallback."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _close(a, b, tol, what):
    err = (a.float() - b.float()).abs().max().item()
    ref = b.float().abs().max().item()
    assert err <= tol * max(ref, 1.0), f"{what}: max err {err:.3e} (ref max {ref:.3e})"


def _sum_over_ranks(t):
    h = t.detach().float().cpu()
    dist.all_reduce(h)
    return h


def _dense_case(rank, world, dev):
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(5)
    N, Ci, Co, H = 4, 32, 64, 32
    x = torch.randn(N, Ci, H, H, generator=g).to(dev).half().contiguous(memory_format=torch.channels_last)
    w = (torch.randn(Co, Ci, 3, 3, generator=g) * 0.06).to(dev)
    gam, bet = (torch.rand(Co, generator=g) + 0.5).to(dev), (torch.randn(Co, generator=g) * 0.1).to(dev)
    gy = torch.randn(N, Co, H, H, generator=g).to(dev).half().contiguous(memory_format=torch.channels_last)
    res = torch.randn(N, Co, H, H, generator=g).to(dev).half().contiguous(memory_format=torch.channels_last)

    def run(bn, xs, rs, gs):
        with torch.no_grad():
            bn.weight.copy_(gam), bn.bias.copy_(bet)
        xs = xs.clone().requires_grad_(True)
        ws = w.clone().requires_grad_(True)
        y = ops.conv_bn_act(xs, ws, bn, True, act="relu", residual=rs)
        y.backward(gs)
        return y.detach(), xs.grad, ws.grad, bn.weight.grad, bn.bias.grad

    full = torch.nn.BatchNorm2d(Co).to(dev)
    y0, dx0, dw0, dg0, db0 = run(full, x, res, gy)
    per = N // world
    sl = slice(rank * per, (rank + 1) * per)
    shard = torch.nn.SyncBatchNorm(Co).to(dev)
    y1, dx1, dw1, dg1, db1 = run(shard, x[sl], res[sl], gy[sl])
    _close(y1, y0[sl], 2e-3, "dense y")
    _close(dx1, dx0[sl], 4e-3, "dense dx")
    _close(_sum_over_ranks(dw1), dw0.cpu(), 4e-3, "dense dw")
    _close(_sum_over_ranks(dg1), dg0.cpu(), 4e-3, "dense dgamma")
    _close(_sum_over_ranks(db1), db0.cpu(), 4e-3, "dense dbeta")
    _close(shard.running_mean, full.running_mean, 1e-4, "running_mean")
    _close(shard.running_var, full.running_var, 1e-4, "running_var")
    assert int(shard.num_batches_tracked) == 1
    # the shard alone (local statistics) must NOT match: the exchange is what makes the difference
    local = torch.nn.BatchNorm2d(Co).to(dev)
    y2 = run(local, x[sl], res[sl], gy[sl])[0]
    assert (y2.float() - y0[sl].float()).abs().max().item() > 1e-2


def _rows_case(rank, world, dev):
    """BatchNorm1d over active sites with DIFFERENT site counts per rank (the count travels with the sums)."""
    from maggie_b200 import ops
    g = torch.Generator().manual_seed(9)
    Ns, Ci, Co = 1000, 64, 32
    src = torch.randn(Ns, Ci, generator=g).to(dev).half()
    w = (torch.randn(Co, 1, 1, Ci, generator=g) * 0.1).to(dev)
    gam, bet = (torch.rand(Co, generator=g) + 0.5).to(dev), (torch.randn(Co, generator=g) * 0.1).to(dev)
    gy = torch.randn(Ns, Co, generator=g).to(dev).half()
    cut = [0, 300, Ns]

    def run(bn, s, gs):
        with torch.no_grad():
            bn.weight.copy_(gam), bn.bias.copy_(bet)
        s = s.clone().requires_grad_(True)
        ws = w.clone().requires_grad_(True)
        y = ops.rows_conv(s, ws, bn=bn, mode="bn_act", act="lrelu", training=True)
        y.backward(gs)
        return y.detach(), s.grad, ws.grad, bn.weight.grad

    full = torch.nn.BatchNorm1d(Co).to(dev)
    y0, ds0, dw0, dg0 = run(full, src, gy)
    sl = slice(cut[rank], cut[rank + 1])
    shard = torch.nn.SyncBatchNorm(Co).to(dev)
    y1, ds1, dw1, dg1 = run(shard, src[sl], gy[sl])
    _close(y1, y0[sl], 2e-3, "rows y")
    _close(ds1, ds0[sl], 4e-3, "rows dsrc")
    _close(_sum_over_ranks(dw1), dw0.cpu(), 4e-3, "rows dw")
    _close(_sum_over_ranks(dg1), dg0.cpu(), 4e-3, "rows dgamma")
    _close(shard.running_var, full.running_var, 1e-4, "rows running_var")


def _model_case(rank, world, dev):
    """Whole model, converted by `convert_sync_batchnorm`, one frame per rank with different content: every rank must end
    up with IDENTICAL running statistics (they come from the exchanged sums), finite losses and gradients."""
    import random
    import numpy as np
    from maggie_b200.config import CfgNode
    from maggie_b200.network import build_model
    from maggie_b200.dp import FlatGradAllReduce
    from oracle import synth
    torch.manual_seed(1234)
    model, _ = build_model(CfgNode(synth.model_cfg()))
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model).to(dev).train()
    model.enable_cuda_graphs(True)           # K15 is captured with the dense stage; the collective fallback runs eagerly
    flat = FlatGradAllReduce(model.parameters())
    batch = synth.make_batch(b=4, n_f=1, n_i=2, H=256, W=256, edge_px=6.0, train=True, it=1)
    batch = {k: (v[2 * rank:2 * rank + 2].to(dev) if torch.is_tensor(v) else v) for k, v in batch.items()}
    np.random.seed(7), random.seed(7)
    flat.zero()
    _, loss = model(batch, mem_feat=None)
    (loss["total"] * 128.0).backward()
    assert torch.isfinite(loss["total"]).item()
    gn = torch.cat([p.grad.flatten() for p in model.parameters() if p.grad is not None]).norm()
    assert torch.isfinite(gn).item() and gn.item() > 0
    for name in ("encoder.bn1", "encoder.layer3.0.bn1", "aspp.aspp5_bn", "decoder.layer1.0.bn2", "decoder.refine_OS1.1"):
        bn = model.get_submodule(name)
        assert isinstance(bn, torch.nn.SyncBatchNorm)
        mine = torch.cat([bn.running_mean, bn.running_var]).cpu()
        both = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(both, mine)
        assert torch.equal(both[0], both[1]), f"{name}: running statistics differ between the ranks"
        assert int(bn.num_batches_tracked) >= 1      # (graph capture adds its warm-up passes)


def _worker(rank, world, port, peer, late_s=0.0):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), MAGGIE_B200_NO_PEER_EXCHANGE="0" if peer else "1")
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.cuda.set_device(0)
        dev = torch.device("cuda:0")
        if late_s:
            # rank 1 reaches its first exchange `late_s` seconds after rank 0 (the reference engine validates on rank 0
            # alone between steps, engine/train.py:294 with `val_dist: false`): the exchange must wait like a collective
            from maggie_b200 import dense
            dense.peer_window(dist.group.WORLD, dev)        # windows are set up together (a collective step)
            if rank == 1:
                import time
                time.sleep(late_s)
            _dense_case(rank, world, dev)
            torch.cuda.synchronize()
            return
        _dense_case(rank, world, dev)
        _rows_case(rank, world, dev)
        from maggie_b200 import dense
        wins = list(dense._WINDOWS.values())
        assert len(wins) == 1 and (wins[0] is not None) == peer, "wrong exchange path"
        if peer:
            _model_case(rank, world, dev)
        torch.cuda.synchronize()
    finally:
        dist.destroy_process_group()


def _spawn(peer, late_s=0.0):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, peer, late_s), nprocs=2, join=True)


def test_peer_memory_exchange_matches_full_batch():
    _spawn(True)


def test_collective_fallback_matches_full_batch():
    _spawn(False)


def test_peer_memory_exchange_waits_for_a_late_rank():
    """One rank shows up 22 s late (longer than the 20 s limit the first version of K15 trapped at): the in-kernel wait
    behaves like the NCCL collective it replaces (watchdog: MAGGIE_B200_XCHG_TIMEOUT_S, default 600 s)."""
    _spawn(True, late_s=22.0)
