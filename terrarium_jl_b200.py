"""Import shim: loads the package that lives in the directory ``terrarium.jl_b200/`` (a dotted
directory name cannot be imported directly) under the module name ``terrarium_jl_b200``."""
import importlib.util
import os
import sys

_root = os.path.join(os.path.dirname(os.path.abspath(__file__)), "terrarium.jl_b200")
_spec = importlib.util.spec_from_file_location(
    "terrarium_jl_b200", os.path.join(_root, "__init__.py"), submodule_search_locations=[_root])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["terrarium_jl_b200"] = _mod
_spec.loader.exec_module(_mod)
