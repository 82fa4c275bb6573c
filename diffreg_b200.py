"""Import alias: the package directory is named ``diff-reg_b200`` (not a valid Python
identifier), so ``import diffreg_b200`` loads it from there."""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg = os.path.join(_here, "diff-reg_b200")
_spec = importlib.util.spec_from_file_location(
    "diffreg_b200", os.path.join(_pkg, "__init__.py"), submodule_search_locations=[_pkg])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["diffreg_b200"] = _mod
_spec.loader.exec_module(_mod)
