"""Minimal stand-in for pymeshfix (see shims/README.md)."""
from . import _meshfix  # noqa: F401
