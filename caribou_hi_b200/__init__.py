"""caribou_hi_b200 -- B200-native (sm_100a CUDA) forward-model likelihood path of
caribou_hi behind the reference's own interface.

The CUDA library (``libchi.so``) is loaded lazily on first use; there is no CPU
fallback anywhere in this package.
"""

from .models import (  # noqa: F401  (same export list as caribou_hi/__init__.py:1-17)
    AbsorptionModel,
    AbsorptionPhysicalModel,
    EmissionAbsorptionModel,
    EmissionAbsorptionPhysicalModel,
    EmissionModel,
    EmissionPhysicalModel,
    HIModel,
    HIPhysicalModel,
)
from .spec_data import SpecData  # noqa: F401
from .engine import Engine, PinnedArray  # noqa: F401
from .optimize import Optimize  # noqa: F401

__version__ = "0.1.0"
