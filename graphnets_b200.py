"""Import shim: the package directory is `graphnets.jl_b200/` (a dot is not importable), so load it
under the module name `graphnets_jl_b200` and re-export its public API.

    import graphnets_b200 as gn
    y = gn.GNBlock((10, 5, 0), (3, 4, 5))(gn.batch(x))
"""
import importlib.util
import os
import sys

_ROOT = os.path.dirname(os.path.abspath(__file__))
_PKG = os.path.join(_ROOT, "graphnets.jl_b200")
_NAME = "graphnets_jl_b200"

if _NAME not in sys.modules:
    _spec = importlib.util.spec_from_file_location(_NAME, os.path.join(_PKG, "__init__.py"),
                                                   submodule_search_locations=[_PKG])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    try:
        _spec.loader.exec_module(_mod)
    except BaseException:
        del sys.modules[_NAME]
        raise
_mod = sys.modules[_NAME]
globals().update({k: getattr(_mod, k) for k in _mod.__all__})
lib = _mod.lib
LIB_PATH = _mod.LIB_PATH
pkg = _mod
__all__ = list(_mod.__all__) + ["lib", "LIB_PATH", "pkg"]
