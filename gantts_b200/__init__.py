"""gantts_b200 -- B200 (sm_100a) native GAN-step hot path of r9y9/gantts behind the reference's
own Python surface (``models`` / ``seqloss`` / ``multistream``).  Host code is Python + PyTorch
(device memory, streams, torch.distributed); all arithmetic on the path runs in hand-written CUDA
through the C ABI of ``libgantts_b200.so`` (include/gantts_b200.h).  No CPU fallback."""
__version__ = "0.1.1+b200"

from . import config  # noqa: F401
from . import models  # noqa: F401
from . import multistream  # noqa: F401
from . import seqloss  # noqa: F401
