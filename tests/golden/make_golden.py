#!/usr/bin/env python3
"""Generate the committed golden fixtures under tests/golden/ from the CPU oracle.

The reference (Rust) cannot be built in this image and ships no numeric golden vectors
for this path, so the fixtures are outputs of the line-faithful oracle
(oracle/bem_oracle.cpp), frozen here so that (a) later edits of the oracle are detected
and (b) the GPU path is compared against bytes that do not depend on the oracle build of
the day.  Regenerate with:  python tests/golden/make_golden.py
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from math_audio_b200.mesh import generate_box_mesh_quad, generate_icosphere_mesh  # noqa: E402
from math_audio_b200.types import PhysicsParams  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OUT = Path(__file__).resolve().parent


def sphere_case(name, sub, ka, a=0.1):
    ph = PhysicsParams.from_wave_number(ka / a)
    mesh = generate_icosphere_mesh(a, sub)
    beta, scale = ph.burton_miller_beta_adaptive(a)
    A, rhs0, nq = orc.assemble(mesh, ph.wave_number, beta)
    rhs, _ = orc.incident_rhs(0, [0, 0, 1.0], 1.0, mesh.center, mesh.normal, ph.wave_number, beta)
    b = rhs0 + rhs
    x, info = orc.gmres(A, b, max_iterations=1000, restart=50, tolerance=1e-10)
    np.savez_compressed(OUT / f"{name}.npz", A=A, b=b, x=x, k=ph.wave_number, beta=beta, sub=sub, a=a,
                        iterations=info["iterations"], restarts=info["restarts"], residual=info["residual"], nqp=nq)
    print(name, A.shape, info)


def box_case(name):
    # small Quad4 "cabinet" with a piston (non-zero, full-length velocity BC) on the front wall
    mesh = generate_box_mesh_quad(0.32, 0.44, 0.64, 4, 6, 8)
    ph = PhysicsParams.new(500.0, 343.0, 1.21, False)
    front = (np.abs(mesh.center[:, 1] + 0.22) < 1e-9) & (np.hypot(mesh.center[:, 0], mesh.center[:, 2]) < 0.12)
    v = np.zeros((mesh.n_elem, 4), dtype=np.complex128)
    v[front] = 1.0
    mesh.set_velocity_bc(v)
    mesh.bc_len[~front] = 1
    beta = ph.burton_miller_beta()
    A, rhs, nq = orc.assemble(mesh, ph.wave_number, beta)
    np.savez_compressed(OUT / f"{name}.npz", A=A, rhs=rhs, k=ph.wave_number, beta=beta, front=front, nqp=nq)
    print(name, A.shape, int(front.sum()), "piston elements")


if __name__ == "__main__":
    sphere_case("ico1_ka0p5", 1, 0.5)    # N=80,  +K' branch boundary (ka*rbar ~ 0.5)
    sphere_case("ico2_ka0p2", 2, 0.2)    # N=320, QA-suite Rayleigh case
    sphere_case("ico2_ka6", 2, 6.0)      # N=320, k*h_e >= 1: exercises the nsec2=3 singular quirk
    box_case("box_4x6x8_piston")
