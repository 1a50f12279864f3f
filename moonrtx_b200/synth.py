"""
Synthetic LOLA-shaped inputs: the 9 GB LDEM and the colour TIFF
cannot be downloaded offline, so tests and bench.py render these instead.

Host (numpy) generators are for the small maps of the tests; the full-size maps
of bench.py (23040x11520, 92160x46080) are generated directly in HBM by the
`mrtx_synth_ldem_i16` / `mrtx_synth_color_bgr` kernels (same recipe: periodic
value-noise fBm + crater bowls, scaled to the real LDEM count range).

Also holds the synthetic ephemeris track that replaces Skyfield for time-lapse
benchmarks (astro.py is out of scope and Skyfield's kernels are not available).
"""

from typing import NamedTuple

import numpy as np

LDEM_MIN_COUNTS = -18200    # -9.1 km at 0.5 m/count (data_loader.py:160-163)
LDEM_MAX_COUNTS = 21600     # +10.8 km


def synth_ldem(W: int, H: int, seed: int = 20240314, craters: int = 400) -> np.ndarray:
    """
    int16 (H, W) equirectangular height map, periodic in longitude: power-law
    (slope ~ -2) fractal relief plus crater bowls with raised rims, scaled to the
    LDEM count range.  FFT based, intended for maps up to ~16 Mpx.
    """
    rng = np.random.default_rng(seed)
    ky = np.fft.fftfreq(H)[:, None] * H
    kx = np.fft.rfftfreq(W)[None, :] * W
    k = np.sqrt(kx * kx + ky * ky)
    k[0, 0] = 1.0
    amp = k ** -1.6
    amp[0, 0] = 0.0
    phase = rng.uniform(0.0, 2.0 * np.pi, size=amp.shape)
    # mirror in latitude so the two poles do not wrap into each other visibly
    relief = np.fft.irfft2(amp * np.exp(1j * phase), s=(H, W))
    relief /= np.abs(relief).max()

    lat = (0.5 - (np.arange(H) + 0.5) / H) * np.pi
    lon = ((np.arange(W) + 0.5) / W - 0.5) * 2.0 * np.pi
    cl, sl = np.cos(lat)[:, None], np.sin(lat)[:, None]
    bowls = np.zeros((H, W), dtype=np.float64)
    n = int(craters)
    c_lat = np.arcsin(rng.uniform(-1, 1, n))
    c_lon = rng.uniform(-np.pi, np.pi, n)
    c_rad = np.radians(0.3 + 8.0 * rng.power(0.35, n))      # angular radius
    c_dep = rng.uniform(0.08, 0.35, n)
    for i in range(n):
        cosd = np.sin(c_lat[i]) * sl + np.cos(c_lat[i]) * cl * np.cos(lon[None, :] - c_lon[i])
        # only touch rows that can be inside 1.6 radii
        rows = np.abs(lat - c_lat[i]) < 1.6 * c_rad[i]
        if not rows.any():
            continue
        d = np.arccos(np.clip(cosd[rows], -1, 1)) / c_rad[i]
        bowl = np.where(d < 1.0, -c_dep[i] * (1.0 - d * d),
                        np.where(d < 1.6, 0.35 * c_dep[i] * np.exp(-((d - 1.0) / 0.25) ** 2), 0.0))
        bowls[rows] += bowl
    z = 0.75 * relief + bowls
    z -= z.min()
    z /= z.max()
    counts = LDEM_MIN_COUNTS + z * (LDEM_MAX_COUNTS - LDEM_MIN_COUNTS)
    return np.rint(counts).astype(np.int16)


def synth_color(W: int, H: int, seed: int = 4720) -> np.ndarray:
    """uint8 BGR (H, W, 3): smooth maria/highland pattern + noise, full 0-255 range."""
    rng = np.random.default_rng(seed)
    y = (np.arange(H) + 0.5)[:, None] / H
    x = (np.arange(W) + 0.5)[None, :] / W
    base = (0.5 + 0.25 * np.sin(2 * np.pi * (3 * x + y)) * np.cos(2 * np.pi * 2 * y)
            + 0.2 * np.sin(2 * np.pi * 7 * x) * np.sin(2 * np.pi * 5 * y))
    out = np.empty((H, W, 3), dtype=np.uint8)
    for c in range(3):
        ch = base * (0.9 + 0.05 * c) + rng.normal(0.0, 0.08, size=(H, W))
        ch -= ch.min()
        ch /= ch.max()
        out[..., c] = np.rint(ch * 255.0).astype(np.uint8)
    return out


class SynthEphemeris(NamedTuple):
    """The fields of the reference's MoonEphemeris (shared_types.py:23-42) that the
    hot path consumes (moon_renderer.py:653-727, 824-860)."""
    distance: float             # km, observer - Moon
    sun_distance: float         # km, Moon - Sun
    phase_angle: float          # deg
    bright_limb_angle: float    # deg
    elongation: float           # deg
    rotation_matrix: np.ndarray


COLONGITUDE_RATE_DEG_PER_HOUR = 0.508   # moon_renderer.py:70


def synth_ephemeris(minutes: float, phase0_deg: float = 90.0) -> SynthEphemeris:
    """
    Terminator sweep of SURVEY.md §8d: sub-observer point (0, 0), bright limb at
    -90 deg (sun to the right), phase angle 90 deg at t=0 (terminator on the central
    meridian) advancing 0.508 deg/h.
    """
    phase = phase0_deg - COLONGITUDE_RATE_DEG_PER_HOUR * minutes / 60.0
    return SynthEphemeris(distance=384_400.0, sun_distance=1.496e8, phase_angle=phase,
                          bright_limb_angle=-90.0, elongation=180.0 - phase,
                          rotation_matrix=np.eye(3))
