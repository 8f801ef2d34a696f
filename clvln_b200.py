"""Import alias for the package directory ``curriculum-learning-for-vln_b200/``.

The directory name the project layout prescribes is not a valid Python identifier, so this
shim loads it as a regular package under the importable name ``clvln_b200``.
"""
import importlib.util
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_pkg_dir = os.path.join(_here, "curriculum-learning-for-vln_b200")
_spec = importlib.util.spec_from_file_location(
    __name__, os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
