"""Alias modules that let the reference's callers run UNCHANGED on this library.

    import agile3d_b200.compat as compat
    compat.install()            # before importing engine.py / eval_multi_obj.py / eval_single_obj.py of the reference

registers
    MinkowskiEngine  -> agile3d_b200.minkowski   (SparseTensor, utils.sparse_quantize, utils.batched_coordinates: the
                        three symbols those callers execute; engine.py:47-51, eval_multi_obj.py:94-98, datasets/*.py)
    models           -> build_model / build_criterion of this package (models/__init__.py:6-10)
and, when asked (seg=True), replaces the click-simulation helpers of utils.seg by their device versions
(agile3d_b200.interactive: same names, same return structure).  matplotlib is imported by evaluation/*.py for plots only;
a stub is registered when it is not installed (stub_missing=True).
"""
from __future__ import annotations

import importlib
import sys
import types


def install(seg: bool = False, stub_missing: bool = True):
    import agile3d_b200
    from . import minkowski

    sys.modules["MinkowskiEngine"] = minkowski
    models = types.ModuleType("models")
    models.build_model = agile3d_b200.build_model
    models.build_criterion = agile3d_b200.build_criterion
    models.__doc__ = "agile3d_b200.compat: models/__init__.py:6-10 of the reference on the B200 library"
    sys.modules["models"] = models
    if stub_missing:
        for name in ("matplotlib", "matplotlib.pyplot"):
            try:
                importlib.import_module(name)
            except ImportError:
                sys.modules[name] = types.ModuleType(name)
        if "matplotlib.pyplot" in sys.modules and "matplotlib" in sys.modules:
            setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])
    if seg:
        from . import interactive
        ref_seg = importlib.import_module("utils.seg")             # the reference's own module (its tree is on sys.path)
        for name in ("get_simulated_clicks", "extend_clicks"):
            setattr(ref_seg, name, getattr(interactive, name))
    return models
