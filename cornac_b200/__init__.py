"""cornac_b200 -- B200 (sm_100a) implementation of Cornac's BPR / MF train-and-rank hot path.

    from cornac_b200 import BPR, MF        # drop-in for cornac.models.BPR / MF
    cornac.Experiment(eval_method=..., models=[BPR(k=64, ...)], metrics=[...]).run()

Layers:  include/b200cornac.h (C ABI)  <-  cornac_b200/csrc (CUDA)  <-  cornac_b200.engine
(ctypes, device tensors)  <-  cornac_b200.recom_bpr / recom_mf (cornac.models.Recommender
plug-ins).  The plug-in classes need the `cornac` package importable (they subclass its
Recommender so that cornac.Experiment accepts them); the engine does not.
"""
__all__ = ["BPR", "WBPR", "MMMF", "VEBPR", "SBPR", "MF", "WMF", "BaselineOnly", "engine", "B200Error"]

from ._lib import B200Error  # noqa: F401


def __getattr__(name):
    if name == "BPR":
        from .recom_bpr import BPR
        return BPR
    if name == "WBPR":
        from .recom_bpr import WBPR
        return WBPR
    if name == "MMMF":
        from .recom_bpr import MMMF
        return MMMF
    if name == "VEBPR":
        from .recom_bprx import VEBPR
        return VEBPR
    if name == "SBPR":
        from .recom_bprx import SBPR
        return SBPR
    if name == "MF":
        from .recom_mf import MF
        return MF
    if name == "WMF":
        from .recom_wmf import WMF
        return WMF
    if name == "BaselineOnly":
        from .recom_bo import BaselineOnly
        return BaselineOnly
    if name == "engine":
        import importlib
        return importlib.import_module(".engine", __name__)
    raise AttributeError(name)
