"""Import shim: the package directory is named `selfconsistentfield.jl_b200` (with a dot,
after the reference repository), which `import` cannot spell.  `import scf_b200` loads that
directory as the package `scf_b200` (submodules: scf_b200.capi, scf_b200.builder, ...)."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "selfconsistentfield.jl_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_dir, "__init__.py"), submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
