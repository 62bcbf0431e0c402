"""Shim package: the reference's top-level `data` namespace (R/data/ has no __init__.py).
Other `data/` directories on sys.path stay importable behind this one."""
from pkgutil import extend_path

__path__ = extend_path(__path__, __name__)
