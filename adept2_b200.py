"""Importable alias of the hyphenated package directory ``adept-2_b200``."""
import importlib as _il
import sys as _sys

_pkg = _il.import_module("adept-2_b200")
_sys.modules[__name__] = _pkg
