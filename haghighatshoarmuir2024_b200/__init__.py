"""haghighatshoarmuir2024_b200 -- B200-native SNN sound-source-localisation hot path.

Drop-in for micloc.snn_beamformer / micloc.spike_encoder / micloc.beamformer
(synsense/HaghighatshoarMuir2024) with the whole STHT -> RZCC -> beamform -> LIF ->
DoA chain running as hand-written sm_100a CUDA behind a C-ABI (include/micloc_b200.h).
"""
from . import _native
from .array_geometry import (ArrayGeometry, CenterCircularArray, CircularArray, LinearArray, Random2DArray)

__all__ = ["ArrayGeometry", "CenterCircularArray", "CircularArray", "LinearArray", "Random2DArray", "_native"]
__version__ = "0.1.0"
