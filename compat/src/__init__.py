"""Shim package: the reference's `src` namespace."""
