"""Import alias for the package directory `regularizedleastsquares.jl_b200/` (its name
contains a dot, which Python's import statement cannot spell)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "regularizedleastsquares.jl_b200")
_spec = importlib.util.spec_from_file_location("rls_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["rls_b200"] = _mod
_spec.loader.exec_module(_mod)
