"""Minimal stand-in for matplotlib (see shims/README.md): plotting calls of the reference scripts become no-ops."""


def use(*args, **kwargs):
    pass
