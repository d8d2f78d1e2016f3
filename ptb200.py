"""Import shim: the package directory `distributed-path-tracer_b200` is not a Python identifier."""
import importlib as _il
import sys as _sys

_pkg = _il.import_module("distributed-path-tracer_b200")
_sys.modules[__name__] = _pkg
