"""Drop-in for the reference's `maggie.network` (network/__init__.py:5-16): `build_model(cfg.model)`."""
import logging
import os

from .arch import ARCHS, MaGGIe, MaGGIe_Temp  # noqa: F401


def build_model(cfg):
    """Returns (model, is_from_hf) like the reference.  Unlike the reference's bare `except: pass` (which leaves
    `model` unbound), a failed hub download raises."""
    arch = ARCHS.get(cfg.arch)
    if arch is None:
        raise NotImplementedError(f"maggie_b200.network implements {sorted(ARCHS)}; got arch={cfg.arch!r}")
    weights = cfg.get("weights", "") if hasattr(cfg, "get") else getattr(cfg, "weights", "")
    if weights != "" and not os.path.isfile(weights):
        model = arch.from_pretrained(weights)
        logging.info(f"Load pretrained model {weights} from Hugging Face")
        return model, True
    return arch(cfg), False
