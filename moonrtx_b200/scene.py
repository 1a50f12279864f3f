"""
Host-side scene arithmetic of the hot path: what MoonRenderer computes per time step
before it touches the renderer (moon_renderer.py:507-544, 653-727, 824-860).  These
are a handful of float64 scalars per frame - they stay on the CPU, as in the reference -
and exist here so that bench.py / the time-lapse driver can produce the per-frame
`(u, v, light_pos, light_radius, eye)` tuples without the Tk application.

Constants are the reference's class constants (moon_renderer.py:36-136).
"""

from typing import NamedTuple

import numpy as np

MOON_RADIUS = 10.0
MOON_RADIUS_KM = 1737.4
MOON_FILL_FRACTION = 0.9
CAMERA_DISTANCE = MOON_RADIUS * 30
MOON_REFERENCE_DISTANCE = 384_400.0
SUN_LIGHT_DISTANCE = 21460
SUN_RADIUS = 100
SUN_RADIUS_KM = 695_700.0
SUN_BRIGHTNESS_SCALE = (2146.0 / 100.0) ** 2
SCENE_EPSILON = 1.0e-4
MARCHING_STEP = 5.0e-3
MARCHING_STEP_EPS = 3.0e-4
ACCUMULATION_FRAMES = 64
PREVIEW_ACCUMULATION_FRAMES = 1
TONEMAP_EXPOSURE = 0.9


class FrameState(NamedTuple):
    """Everything that changes between two time steps (moon_renderer.py:840-860)."""
    u: tuple            # scene direction of the body north pole  (R[:, 2])
    v: tuple            # scene direction of longitude 0          (-R[:, 1])
    light_pos: tuple
    light_radius: float
    eye: tuple
    target: tuple
    up: tuple
    fov: float


def moon_apparent_radius(distance_km: float) -> float:
    """moon_renderer.py:522-529"""
    return float(np.arcsin(MOON_RADIUS_KM / distance_km))


def moon_camera_distance(distance_km: float) -> float:
    """moon_renderer.py:531-544: eye distance that shows the Moon at its true apparent size."""
    return CAMERA_DISTANCE * (moon_apparent_radius(MOON_REFERENCE_DISTANCE) / moon_apparent_radius(distance_km))


def default_fov() -> float:
    """moon_renderer.py:513-520: vertical fov of the whole-disk view, 4.2422 deg."""
    visible_height = 2 * MOON_RADIUS / MOON_FILL_FRACTION
    fov = np.degrees(2 * np.arctan(visible_height / (2 * CAMERA_DISTANCE)))
    return float(max(1, min(90, fov)))


def light_position(bright_limb_angle_deg: float, phase_angle_deg: float) -> tuple:
    """moon_renderer.py:678-727: sun direction from the bright-limb and phase angles."""
    b = np.radians(bright_limb_angle_deg)
    p = np.radians(phase_angle_deg)
    d = SUN_LIGHT_DISTANCE
    return (float(-np.sin(b) * np.sin(p) * d), float(-np.cos(p) * d), float(np.cos(b) * np.sin(p) * d))


def light_radius(sun_distance_km: float) -> float:
    """moon_renderer.py:859: light radius keeps the true solar angular size."""
    return float(SUN_LIGHT_DISTANCE * SUN_RADIUS_KM / sun_distance_km)


def light_radiance(brightness: float) -> float:
    """moon_renderer.py:640: the light 'color' is a radiance."""
    return float(brightness * SUN_BRIGHTNESS_SCALE)


def frame_state(ephem, eye_direction=(0.0, -1.0, 0.0)) -> FrameState:
    """
    The scene of one time step from an ephemeris record with the fields of the
    reference's MoonEphemeris (distance, sun_distance, phase_angle, bright_limb_angle,
    rotation_matrix): update_view, moon_renderer.py:840-860, with the default camera.
    """
    R = np.asarray(ephem.rotation_matrix, dtype=np.float64)
    dist = moon_camera_distance(ephem.distance)
    e = np.asarray(eye_direction, dtype=np.float64)
    e = e / np.linalg.norm(e) * dist
    return FrameState(
        u=tuple(float(x) for x in R[:, 2]),
        v=tuple(float(-x) for x in R[:, 1]),
        light_pos=light_position(ephem.bright_limb_angle, ephem.phase_angle),
        light_radius=light_radius(ephem.sun_distance),
        eye=tuple(float(x) for x in e), target=(0.0, 0.0, 0.0), up=(0.0, 0.0, 1.0),
        fov=default_fov(),
    )
