"""pylinac_b200 -- B200-native (sm_100a CUDA, ctypes C-ABI) implementation of pylinac's 2-D EPID image hot path.

Drop-in for the path ``core.image`` / ``core.profile`` primitives -> ``PicketFence`` / ``WinstonLutz`` / ``Starshot`` /
``FieldAnalysis`` ``.analyze()``; see DESIGN.md.  There is no CPU fallback: compute calls need a CUDA device.
"""
from .version import __version__  # noqa: F401
