"""deepdrr_b200 -- B200-native (sm_100a) DRR projection path behind the ``deepdrr.Projector`` API.

Only the projection hot path of arcadelab/deepdrr lives here: the Projector, the
volume / material / spectrum data contracts it consumes, and the CUDA library behind it.
"""
from . import geo, vol
from .material import Material
from .projector import DeprecationError, Projector
from .vol import HUVolume, Mesh, Volume

__all__ = ["Projector", "Volume", "HUVolume", "Mesh", "Material", "geo", "vol", "DeprecationError"]
