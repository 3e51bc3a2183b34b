"""-m gpu: the model under the reference ENGINE's training-loop calls, as `tools/main.py` / `engine/train.py` drive it
(engine/train.py:150-164: build_model -> .to(device) -> SyncBatchNorm.convert_sync_batchnorm -> DistributedDataParallel(
find_unused_parameters); :222-231: optimizer.zero_grad, autocast, model(batch, mem_feat=None); :265-283: GradScaler
scale / unscale_ / clip_grad_norm_(all, 0.01) / step / update, lr_scheduler.step), plus the checkpoint round trips the
engine performs (state_dict strict, save_pretrained / from_pretrained).  The reference sources are not needed here; the
same drop-in is exercised against the reference's own modules by tests/test_reference_engine_api.py in the build container."""
import itertools
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist

from maggie_b200.config import CfgNode
from maggie_b200.network import build_model
from oracle import make_golden as G
from oracle import synth

pytestmark = pytest.mark.gpu


def _dev(batch):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in batch.items()}


@pytest.fixture()
def one_rank_nccl():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(0)
    dist.init_process_group("nccl", rank=0, world_size=1, device_id=torch.device("cuda:0"))
    yield
    dist.destroy_process_group()


def test_reference_training_loop_calls(one_rank_nccl):
    device = torch.device("cuda:0")
    torch.manual_seed(1234)
    model, is_from_hf = build_model(CfgNode(synth.model_cfg()))                      # train.py:150
    model = model.to(device)                                                         # :151
    optimizer = torch.optim.AdamW(model.parameters(), lr=1.5e-4, betas=(0.9, 0.999), weight_decay=0.01)   # optim.py:118
    lr_scheduler = torch.optim.lr_scheduler.OneCycleLR(optimizer, max_lr=1.5e-4, total_steps=100, pct_start=0.1,
                                                       anneal_strategy="cos", cycle_momentum=False)       # optim.py:137
    model = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)                     # :160-161
    model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[0], find_unused_parameters=True)  # :163-164
    model.train()
    scaler = torch.amp.GradScaler("cuda")                                            # :208
    before = {n: p.detach().clone() for n, p in model.module.named_parameters()}
    losses = []
    for it in (1, 2):
        batch = _dev(synth.make_batch(b=2, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=40 + it, train=True, it=it))
        batch = {k: v for k, v in batch.items() if k not in ("fg", "bg")}
        batch["iter"] = it
        G.seed_all(it)
        optimizer.zero_grad()                                                        # :224
        with torch.autocast("cuda"):                                                 # :226
            output, loss = model(batch, mem_feat=None)                               # :227
        scaler.scale(loss["total"]).backward()                                       # :266
        scaler.unscale_(optimizer)                                                   # :272
        all_params = itertools.chain(*[x["params"] for x in optimizer.param_groups])
        norm = torch.nn.utils.clip_grad_norm_(all_params, 0.01)                      # :273-274
        scaler.step(optimizer)                                                       # :278
        scaler.update()
        lr_scheduler.step()                                                          # :283
        assert torch.isfinite(norm) and float(norm) > 0
        losses.append(float(loss["total"]))
        assert output["refined_masks"].shape == (2, 1, 2, 128, 128)
    assert all(np.isfinite(losses))
    assert scaler.get_scale() >= 65536.0, "GradScaler backed off: a step saw inf / nan gradients"
    moved = [n for n, p in model.module.named_parameters() if not torch.equal(p, before[n])]
    frozen = [n for n, p in model.module.named_parameters() if torch.equal(p, before[n]) and p.requires_grad]
    assert len(moved) > 280
    # only parameters outside the graph stay put (dummy_downscale: trainable flag, no gradient - hence find_unused_parameters)
    assert all(("dummy_downscale" in n) or n.endswith(("weight_u", "weight_v")) for n in frozen), frozen[:5]
    assert any(isinstance(m, torch.nn.SyncBatchNorm) for m in model.modules())


def test_checkpoint_round_trips(tmp_path):
    m, _ = build_model(CfgNode(synth.model_cfg()))
    m.load_state_dict(synth.synth_state_dict(m.state_dict()), strict=True)
    m = m.cuda().eval()
    batch = _dev(synth.make_batch(b=1, n_f=1, n_i=2, H=128, W=128, edge_px=4.0, seed=5))
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    with torch.no_grad():
        ref = m(batch, mem_feat=None)["alpha_os8"].clone()
    # engine checkpoint: torch.save(model.state_dict()) / load_state_dict(strict=True)  (train.py:171-185, 310-316)
    path = tmp_path / "last_model.pth"
    torch.save(sd, path)
    m2, _ = build_model(CfgNode(synth.model_cfg()))
    missing, unexpected = m2.load_state_dict(torch.load(path), strict=True)
    assert not missing and not unexpected
    m2 = m2.cuda().eval()
    with torch.no_grad():
        out = m2(batch, mem_feat=None)["alpha_os8"]
    assert torch.equal(out, ref)                        # same weights, same spectral-norm state: bit-identical eval
    # Hugging Face hub mixin (arch/maggie.py:18, the route `model.weights` of the configs takes)
    if hasattr(m2, "save_pretrained"):
        m2.load_state_dict(sd)
        m2.save_pretrained(tmp_path / "hf")
        m3 = type(m2).from_pretrained(tmp_path / "hf", cfg=CfgNode(synth.model_cfg()))
        got, want = m3.state_dict(), sd
        assert set(got) == set(want) and all(torch.equal(got[k].cpu(), want[k].cpu()) for k in want)
