"""Process-wide knobs of the B200 path."""
import os

from . import _lib

_ENGINES = {"simt": _lib.ENGINE_SIMT, "tc": _lib.ENGINE_TC}

# GEMM engine for the fused Linear layers: "tc" = tcgen05 tensor cores (bf16x3 split, fp32
# accumulation in TMEM), "simt" = exact-fp32 FFMA tiles (validation path).
engine = os.environ.get("GANTTS_B200_ENGINE", "tc").lower()


def engine_id(name=None):
    name = engine if name is None else name
    if name not in _ENGINES:
        raise RuntimeError("gantts_b200: unknown engine %r (use 'tc' or 'simt')" % (name,))
    return _ENGINES[name]
