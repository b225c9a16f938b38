"""Import alias: registers the package directory ``once-for-both_b200/`` as the importable package ``ofb_b200``."""
import importlib.util
import os
import sys

_pkg_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "once-for-both_b200")
_spec = importlib.util.spec_from_file_location(
    "ofb_b200", os.path.join(_pkg_dir, "__init__.py"), submodule_search_locations=[_pkg_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["ofb_b200"] = _mod
_spec.loader.exec_module(_mod)
