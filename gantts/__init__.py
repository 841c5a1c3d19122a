"""``gantts`` -- drop-in alias of the B200-native implementation (``gantts_b200``), so that the
reference's ``train.py`` / ``evaluation_*.py`` (``import gantts``, ``from gantts.models import ...``,
``getattr(gantts.models, hp.generator)``; train.py:53-57,773-774) pick up the CUDA path unchanged."""
import sys

import gantts_b200
from gantts_b200 import models, multistream, seqloss  # noqa: F401

__version__ = "0.1.1"

sys.modules[__name__ + ".models"] = models
sys.modules[__name__ + ".multistream"] = multistream
sys.modules[__name__ + ".seqloss"] = seqloss
