"""Import shim: ``import dzopt_b200`` loads the package directory ``dzoptimization.jl_b200/``
(whose name, fixed by the project layout, contains a dot and so cannot be imported by name)."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "dzoptimization.jl_b200")
_NAME = "dzoptimization_jl_b200"

if _NAME in sys.modules:
    _mod = sys.modules[_NAME]
else:
    _spec = importlib.util.spec_from_file_location(
        _NAME, os.path.join(_PKG_DIR, "__init__.py"), submodule_search_locations=[_PKG_DIR])
    _mod = importlib.util.module_from_spec(_spec)
    sys.modules[_NAME] = _mod
    _spec.loader.exec_module(_mod)
sys.modules[__name__] = _mod
