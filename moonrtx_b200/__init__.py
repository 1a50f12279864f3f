"""
moonrtx_b200 - B200-native replacement for the data-parallel hot path of MoonRTX
(albireo77/moonrtx): LDEM/colour downscale, displaced-sphere primary + sun-shadow ray
tracing, shading, progressive accumulation, tone mapping, frame-parallel time-lapse.

Hand-written sm_100a CUDA behind a C ABI (include/moonb200.h), bound with ctypes.
No CPU fallback: importing works anywhere, computing needs libmoonb200.so and a B200.
"""

__version__ = "0.1.0"
