"""Host-side mirror of ``math-bem/src/core/postprocess/pressure.rs``: evaluation-point generators
(:311-424), ``FieldPoint`` (:24-56), ``compute_scattered_field`` (:81-137), ``compute_total_field``
(:273-309) and ``compute_rcs`` (:438-478).  The O(M N) sums run on the device
(``csrc/postprocess.cu`` behind ``bemb200_scattered_field`` / ``bemb200_compute_rcs``)."""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List, Optional

import numpy as np

from . import bem
from .incident import IncidentField
from .types import PhysicsParams


@dataclass
class FieldPoint:
    """pressure.rs:24-56."""
    position: np.ndarray
    p_incident: complex
    p_scattered: complex

    @property
    def p_total(self) -> complex:
        return self.p_incident + self.p_scattered

    def magnitude(self) -> float:
        return abs(self.p_total)

    def spl_db(self) -> float:
        return 20.0 * math.log10(abs(self.p_total) / 20e-6)


def generate_sphere_eval_points(radius: float, n_theta: int, n_phi: int) -> np.ndarray:
    """pressure.rs:311-330: theta at cell centres, phi from 0."""
    pts = np.empty((n_theta * n_phi, 3))
    r = 0
    for i in range(n_theta):
        theta = math.pi * (i + 0.5) / n_theta
        st, ct = math.sin(theta), math.cos(theta)
        for j in range(n_phi):
            phi = 2.0 * math.pi * j / n_phi
            pts[r] = (radius * st * math.cos(phi), radius * st * math.sin(phi), radius * ct)
            r += 1
    return pts


def generate_line_eval_points(start, end, n_points: int) -> np.ndarray:
    """pressure.rs:332-343."""
    pts = np.empty((n_points, 3))
    den = max(n_points - 1, 1)
    for i in range(n_points):
        t = i / den
        pts[i] = [start[d] + t * (end[d] - start[d]) for d in range(3)]
    return pts


def generate_plane_eval_points(center, normal, extent: float, n_points: int) -> np.ndarray:
    """pressure.rs:345-424: n_points x n_points grid spanning [-extent, extent]^2 in the plane through ``center``."""
    n = np.asarray(normal, dtype=np.float64)
    n = n / math.sqrt(float(n @ n))
    arbitrary = np.array([1.0, 0.0, 0.0]) if abs(n[0]) < 0.9 else np.array([0.0, 1.0, 0.0])
    u = np.array([n[1] * arbitrary[2] - n[2] * arbitrary[1], n[2] * arbitrary[0] - n[0] * arbitrary[2], n[0] * arbitrary[1] - n[1] * arbitrary[0]])
    u = u / math.sqrt(float(u @ u))
    v = np.array([n[1] * u[2] - n[2] * u[1], n[2] * u[0] - n[0] * u[2], n[0] * u[1] - n[1] * u[0]])
    den = max(n_points - 1, 1)
    pts = np.empty((n_points * n_points, 3))
    r = 0
    for i in range(n_points):
        s = -extent + 2.0 * extent * i / den
        for j in range(n_points):
            t = -extent + 2.0 * extent * j / den
            pts[r] = [center[d] + s * u[d] + t * v[d] for d in range(3)]
            r += 1
    return pts


def compute_scattered_field(eval_points, staged: bem.StagedMesh, surface_pressure, surface_velocity, physics: PhysicsParams) -> np.ndarray:
    """pressure.rs:81-137 (device)."""
    return bem.compute_scattered_field(eval_points, staged, surface_pressure, surface_velocity, physics)


def compute_total_field(eval_points, staged: bem.StagedMesh, surface_pressure, surface_velocity: Optional[np.ndarray],
                        incident_field: IncidentField, physics: PhysicsParams) -> List[FieldPoint]:
    """pressure.rs:273-309: incident (host, O(M)) + scattered (device, O(M N))."""
    pts = np.ascontiguousarray(eval_points, dtype=np.float64).reshape(-1, 3)
    p_inc = incident_field.evaluate_pressure(pts, physics)
    p_sc = bem.compute_scattered_field(pts, staged, surface_pressure, surface_velocity, physics)
    return [FieldPoint(pts[i].copy(), complex(p_inc[i]), complex(p_sc[i])) for i in range(pts.shape[0])]


def compute_rcs(surface_pressure, staged: bem.StagedMesh, direction, physics: PhysicsParams):
    """pressure.rs:438-478 (device)."""
    return bem.compute_rcs(surface_pressure, staged, direction, physics)
