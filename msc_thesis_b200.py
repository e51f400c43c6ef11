"""Importable alias of the package directory `msc-thesis_b200/` (a hyphen cannot appear in an import
statement): `import msc_thesis_b200 as h3`."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("msc-thesis_b200")
sys.modules[__name__] = _pkg
