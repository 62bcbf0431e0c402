"""Shim package: the reference's top-level `data` namespace."""
