"""Build container only (skipped where /root/reference is absent): the package is a drop-in for `maggie.network` under the
reference's OWN engine modules - `maggie.engine.train` imports `build_model` from it, and the reference's
`build_optim_lr_scheduler` (engine/optim.py:97-140, imported unchanged) builds its optimizer / scheduler on the model."""
import os
import sys
import types

import pytest
import torch

from oracle import ref_shims, synth

pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(ref_shims.REFERENCE_ROOT, "maggie")),
                                reason="needs the reference checkout (build container)")


@pytest.fixture()
def reference_engine(monkeypatch):
    ref_shims.install()
    import maggie_b200.network as ours
    # the one-line switch of INTEGRATION.md: `maggie.network` is this package's network module
    import maggie                                            # the reference package (namespace only)
    monkeypatch.setitem(sys.modules, "maggie.network", ours)
    monkeypatch.setattr(maggie, "network", ours, raising=False)
    # third-party / data modules the engine imports at module level and this image lacks
    for name, attrs in (("wandb", {"Image": object, "log": lambda *a, **k: None}),
                        ("maggie.dataloader", {"build_dataset": lambda *a, **k: None}),
                        ("maggie.utils.metric", {"build_metric": lambda *a, **k: None}),
                        ("maggie.engine.test", {"eval_image": None, "eval_video": None, "test": None})):
        if name not in sys.modules:
            mod = types.ModuleType(name)
            for k, v in attrs.items():
                setattr(mod, k, v)
            monkeypatch.setitem(sys.modules, name, mod)
    for name in ("maggie.engine", "maggie.engine.train", "maggie.engine.optim"):
        monkeypatch.delitem(sys.modules, name, raising=False)
    import importlib
    train = importlib.import_module("maggie.engine.train")
    optim = importlib.import_module("maggie.engine.optim")
    return ours, train, optim


def test_engine_builds_this_model_and_its_optimizer(reference_engine):
    ours, train, optim = reference_engine
    assert train.build_model is ours.build_model             # engine/train.py:13 now resolves to this package
    cfg = ref_shims.CfgNode(dict(
        model=synth.model_cfg(),
        train=dict(max_iter=52000,
                   optimizer=dict(name="adamw", lr=1.5e-4, betas=[0.9, 0.999], momentum=0.9, weight_decay=0.01),
                   scheduler=dict(name="cosine", warmup_iters=1000, gamma=0.1, power=0.9, step_size=10000))))
    model, is_from_hf = train.build_model(cfg.model)          # engine/train.py:150
    assert not is_from_hf and len(model.state_dict()) == 619
    optimizer, scheduler = optim.build_optim_lr_scheduler(cfg, model)      # engine/train.py:156 (reference code, unchanged)
    assert isinstance(optimizer, torch.optim.AdamW)
    n_opt = sum(p.numel() for g in optimizer.param_groups for p in g["params"])
    assert n_opt == sum(p.numel() for p in model.parameters())
    # the engine's SyncBatchNorm conversion keeps the state-dict contract
    conv = torch.nn.SyncBatchNorm.convert_sync_batchnorm(model)
    assert set(conv.state_dict()) == set(model.state_dict())
    lr0 = scheduler.get_last_lr()[0]
    optimizer.step()
    scheduler.step()
    assert scheduler.get_last_lr()[0] > lr0                   # OneCycle warm-up of the reference config
