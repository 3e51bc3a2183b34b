from .maggie import MaGGIe, MaGGIe_Temp  # noqa: F401

ARCHS = {"MaGGIe": MaGGIe, "MaGGIe_Temp": MaGGIe_Temp}
