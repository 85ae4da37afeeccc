"""Import shim: the package directory is literally `muscle.jl_b200/` (not an importable name), so
`import muscle_b200` loads it from there and registers it under this module name."""
import importlib.util as _u
import os as _os
import sys as _sys

_d = _os.path.join(_os.path.dirname(_os.path.abspath(__file__)), "muscle.jl_b200")
_spec = _u.spec_from_file_location("muscle_b200", _os.path.join(_d, "__init__.py"),
                                   submodule_search_locations=[_d])
_mod = _u.module_from_spec(_spec)
_sys.modules["muscle_b200"] = _mod
_spec.loader.exec_module(_mod)
