"""Shim: only ``tf.contrib.training.HParams`` (reference hparams.py:3,16) is provided."""
from . import contrib  # noqa: F401
