"""nnmnkwii.autograd shim: ``unit_variance_mlpg(R, means)`` -> the CUDA stencil+FIR MLPG op."""
from gantts_b200.ops import unit_variance_mlpg  # noqa: F401
