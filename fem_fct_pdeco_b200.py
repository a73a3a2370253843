"""Importable alias of the package directory ``fem-fct-pdeco_b200/`` (a hyphen cannot appear in a Python
module name).  ``import fem_fct_pdeco_b200`` executes this file, which loads the real package from the
hyphenated directory under the same module name and replaces itself in ``sys.modules``."""
import importlib.util
import os
import sys

_dir = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fem-fct-pdeco_b200")
_spec = importlib.util.spec_from_file_location(__name__, os.path.join(_dir, "__init__.py"),
                                               submodule_search_locations=[_dir])
_mod = importlib.util.module_from_spec(_spec)
sys.modules[__name__] = _mod
_spec.loader.exec_module(_mod)
