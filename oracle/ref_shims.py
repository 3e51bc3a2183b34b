"""sys.modules shims that let the UNMODIFIED reference (`/root/reference/maggie`) import in this image.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Only usable where /root/reference exists (the build
container); the GPU box never calls this.  Missing third-party packages and what stands in:
  yacs.config.CfgNode        -> attribute dict (maggie/network/arch/maggie.py:10,21-22)
  kornia.morphology.dilation -> import-only stub (maggie/utils/utils.py:5; the call is commented out)
  fvcore.nn.weight_init      -> import-only stub (maggie/network/module/mask_attention.py:1, dead class)
  spconv.pytorch             -> oracle/spconv_torch.py
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MAGGIE_REFERENCE_ROOT", "/root/reference")


class CfgNode(dict):
    """Minimal yacs.config.CfgNode: nested attribute-access dict."""

    def __init__(self, init_dict=None, **_):
        super().__init__()
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def get(self, k, default=None):
        return self[k] if k in self else default


def install():
    """Register the stub modules and put the reference on sys.path. Idempotent."""
    if "yacs" not in sys.modules:
        yacs = types.ModuleType("yacs")
        yacs_config = types.ModuleType("yacs.config")
        yacs_config.CfgNode = CfgNode
        yacs.config = yacs_config
        sys.modules["yacs"], sys.modules["yacs.config"] = yacs, yacs_config
    if "kornia" not in sys.modules:
        kornia = types.ModuleType("kornia")
        morph = types.ModuleType("kornia.morphology")
        morph.dilation = lambda *a, **k: (_ for _ in ()).throw(NotImplementedError("kornia stub"))
        kornia.morphology = morph
        sys.modules["kornia"], sys.modules["kornia.morphology"] = kornia, morph
    if "fvcore" not in sys.modules:
        fvcore = types.ModuleType("fvcore")
        fnn = types.ModuleType("fvcore.nn")
        wi = types.ModuleType("fvcore.nn.weight_init")
        wi.c2_xavier_fill = lambda m: None
        fnn.weight_init = wi
        fvcore.nn = fnn
        sys.modules.update({"fvcore": fvcore, "fvcore.nn": fnn, "fvcore.nn.weight_init": wi})
    if "spconv" not in sys.modules:
        from . import spconv_torch

        spconv = types.ModuleType("spconv")
        spconv.pytorch = spconv_torch
        sys.modules["spconv"], sys.modules["spconv.pytorch"] = spconv, spconv_torch
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "maggie", "network"))


def import_reference_network():
    """Returns the reference's `maggie.network` package (unmodified source)."""
    install()
    import maggie.network as net  # noqa: E402  (from /root/reference)

    return net
