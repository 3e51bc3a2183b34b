"""Minimal attribute-dict config node (the reference uses yacs.CfgNode, which is absent from this image;
arch/maggie.py:21-22 also accepts a plain dict, which is what `from_pretrained` passes)."""


class CfgNode(dict):
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


def as_cfg(cfg):
    """dict / yacs CfgNode / our CfgNode -> our CfgNode (deep copy of the mapping structure)."""
    if isinstance(cfg, CfgNode):
        return cfg
    return CfgNode({k: (dict(v) if hasattr(v, "keys") else v) for k, v in dict(cfg).items()})
