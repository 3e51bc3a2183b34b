from .maggie import MaGGIe  # noqa: F401

ARCHS = {"MaGGIe": MaGGIe}
