"""PhysicsParams and the Burton-Miller coupling variants.

Host-side mirror of math-bem/src/core/types.rs:16-219 (same names, same
thresholds).  Pure scalar host logic; nothing here touches the device.
"""
from __future__ import annotations

import math
from dataclasses import dataclass


@dataclass
class PhysicsParams:
    """types.rs:16-58.  ``harmonic_factor`` = +1  =>  exp(+ikr)."""

    speed_of_sound: float
    density: float
    frequency: float
    wave_number: float
    omega: float
    wave_length: float
    harmonic_factor: float
    pressure_factor: float
    tau: float

    @staticmethod
    def new(frequency: float, speed_of_sound: float, density: float, is_internal: bool) -> "PhysicsParams":
        omega = 2.0 * math.pi * frequency
        k = omega / speed_of_sound
        hf = 1.0
        return PhysicsParams(
            speed_of_sound=speed_of_sound,
            density=density,
            frequency=frequency,
            wave_number=k,
            omega=omega,
            wave_length=speed_of_sound / frequency,
            harmonic_factor=hf,
            pressure_factor=density * omega * hf,
            tau=-1.0 if is_internal else 1.0,
        )

    @staticmethod
    def from_wave_number(k: float, speed_of_sound: float = 343.0, density: float = 1.21) -> "PhysicsParams":
        """qa_suite.rs:210-212: freq = k c / (2 pi); then PhysicsParams::new."""
        return PhysicsParams.new(k * speed_of_sound / (2.0 * math.pi), speed_of_sound, density, False)

    def gamma(self) -> float:  # types.rs:216-218
        return 1.0

    def burton_miller_beta(self) -> complex:  # types.rs:64-70
        if self.tau > 0.0:
            return complex(0.0, self.harmonic_factor / self.wave_number)
        return 0j

    def burton_miller_beta_bounded(self, k_ref: float) -> complex:  # types.rs:81-87
        if self.tau > 0.0:
            return complex(0.0, self.harmonic_factor / (self.wave_number + k_ref))
        return 0j

    def burton_miller_beta_floored(self, edge_e_magnitude: float, min_beta_e: float) -> complex:  # types.rs:100-117
        if self.tau > 0.0:
            eta = max(1.0 / self.wave_number, min_beta_e / edge_e_magnitude)
            return complex(0.0, self.harmonic_factor * eta)
        return 0j

    @staticmethod
    def optimal_beta_scale(ka: float) -> float:  # types.rs:201-213
        if ka < 0.85:
            return 32.0
        if ka < 0.92:
            return 8.0
        if ka < 1.2:
            return 4.0
        if ka < 1.8:
            return 8.0
        return 16.0

    def burton_miller_beta_optimal(self, element_size: float) -> complex:  # types.rs:124-131
        if self.tau > 0.0:
            k_ref = 1.0 / element_size
            return complex(0.0, self.harmonic_factor / (self.wave_number + k_ref))
        return 0j

    def burton_miller_beta_scaled(self, scale: float) -> complex:  # types.rs:144-150
        if self.tau > 0.0:
            return complex(0.0, self.harmonic_factor * scale / self.wave_number)
        return 0j

    def burton_miller_beta_adaptive(self, radius: float):  # types.rs:173-195
        if self.tau <= 0.0:
            return 0j, 1.0
        ka = self.wave_number * radius
        if ka < 0.5:
            scale = 1.0
        elif ka < 1.2:
            scale = 4.0
        elif ka < 1.8:
            scale = 8.0
        else:
            scale = 16.0
        return complex(0.0, self.harmonic_factor * scale / self.wave_number), scale
