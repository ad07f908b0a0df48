"""Array geometries: per-microphone propagation delays for a plane wave.

Host-side mirror of the reference's micloc/array_geometry.py (same class names,
constructor arguments and `delays` semantics, file:line cited per class); used to
synthesise inputs and to design beamforming matrices.  Not on the device path.
"""
from __future__ import annotations

import numpy as np

SOUND_SPEED_IN_OPEN_AIR = 340


class ArrayGeometry:
    """Polar sensor positions -> relative delays (micloc/array_geometry.py:17-60)."""

    def __init__(self, r_vec: np.ndarray, theta_vec: np.ndarray, speed: float = SOUND_SPEED_IN_OPEN_AIR):
        r_vec = np.asarray(r_vec, dtype=np.float64)
        if np.any(r_vec < 0):
            raise ValueError("distances of the elements in `r_vec` should be all positive!")
        self.r_vec = r_vec
        self.theta_vec = np.asarray(theta_vec, dtype=np.float64)
        self.speed = speed

    def delays(self, theta: float, normalized: bool = True) -> np.ndarray:
        """delay_m = -r_m cos(theta_m - theta) / speed, optionally shifted to start at 0."""
        d = -self.r_vec * np.cos(self.theta_vec - theta) / self.speed
        if normalized:
            d = d - np.min(d)
        return d

    def delays_batch(self, theta: np.ndarray) -> np.ndarray:
        """Un-normalised delays for many DoAs at once: [len(theta), M]."""
        theta = np.asarray(theta, dtype=np.float64).reshape(-1, 1)
        return -self.r_vec[None, :] * np.cos(self.theta_vec[None, :] - theta) / self.speed

    def __len__(self) -> int:
        return len(self.r_vec)


class CircularArray(ArrayGeometry):
    """num_mic sensors on a circle, angles linspace(0, 2pi, num_mic) (array_geometry.py:63-78)."""

    def __init__(self, radius: float, num_mic: int, speed: float = SOUND_SPEED_IN_OPEN_AIR):
        super().__init__(radius * np.ones(num_mic), np.linspace(0, 2 * np.pi, num_mic), speed)


class CenterCircularArray(ArrayGeometry):
    """num_mic-1 sensors on the circle plus one at the centre (array_geometry.py:81-95).

    As upstream, the ring angles are linspace(0, 2pi, num_mic-1), so the first and
    last ring sensors coincide."""

    def __init__(self, radius: float, num_mic: int, speed: float = SOUND_SPEED_IN_OPEN_AIR):
        r_vec = np.concatenate([radius * np.ones(num_mic - 1), [0.0]])
        theta_vec = np.concatenate([np.linspace(0, 2 * np.pi, num_mic - 1), [0.0]])
        super().__init__(r_vec, theta_vec, speed)


class LinearArray(ArrayGeometry):
    """Uniform line centred at the origin (array_geometry.py:98-120)."""

    def __init__(self, spacing: float, num_mic: int, radius: float, speed: float = SOUND_SPEED_IN_OPEN_AIR):
        x = spacing * (np.arange(-num_mic / 2, num_mic / 2) + 0.5)
        theta_vec = np.where(x < 0, np.pi, 0.0)
        super().__init__(np.abs(x), theta_vec, speed)
        self.radius = radius


class Random2DArray(ArrayGeometry):
    """Sensors uniform in a disc, drawn from numpy's global RNG (array_geometry.py:123-131)."""

    def __init__(self, radius: float, num_mic: int, speed: float = SOUND_SPEED_IN_OPEN_AIR):
        r_vec = np.sqrt(np.random.rand(num_mic)) * radius
        theta_vec = np.random.rand(num_mic) * 2 * np.pi
        super().__init__(r_vec, theta_vec, speed)
        self.radius = radius
