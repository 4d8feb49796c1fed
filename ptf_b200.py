"""Import shim: the product package lives in ``passivetracerflows.jl_b200/`` (a directory name Python cannot
import directly because of the dot).  ``import ptf_b200`` gives that package."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "passivetracerflows.jl_b200")
_spec = importlib.util.spec_from_file_location("ptf_b200", os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ptf_b200"] = _mod
_spec.loader.exec_module(_mod)
