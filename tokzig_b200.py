"""Import shim: the package directory is named ``tokenizer-zig_b200`` (not an importable identifier), so it is loaded
here under the module name ``tokzig_b200``.  ``import tokzig_b200`` from the repo root gives the package."""
import importlib.util
import os
import sys

_PKG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "tokenizer-zig_b200")
_spec = importlib.util.spec_from_file_location("tokzig_b200", os.path.join(_PKG_DIR, "__init__.py"),
                                               submodule_search_locations=[_PKG_DIR])
_mod = importlib.util.module_from_spec(_spec)
sys.modules["tokzig_b200"] = _mod
_spec.loader.exec_module(_mod)
