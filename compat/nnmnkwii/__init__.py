"""Shim of the nnmnkwii symbols the reference imports (see compat/README.md)."""
from . import paramgen, autograd, preprocessing, metrics, datasets  # noqa: F401
