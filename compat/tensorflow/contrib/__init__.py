from . import training  # noqa: F401
