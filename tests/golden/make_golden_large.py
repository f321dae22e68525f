#!/usr/bin/env python3
"""Golden fixtures for the LARGE configurations (BASELINE.json configs[1] and configs[3]), generated once
by the CPU oracle (oracle/bem_oracle.cpp) and committed, so that neither `bench.py` nor the `-m gpu`
tests need the oracle at these sizes at run time:

  config4_rows.npz   121 680-element geodesic sphere (nu = 78), ka = 16, adaptive beta: 32 sampled rows
                     (4 per rank of an 8-way row partition: first/last row of the slab and two interior
                     ones), each at the 256 columns nearest to the collocation point (all subdivided
                     pairs, the self term) plus every 32nd column, and the dot product of the WHOLE row
                     with a seeded vector (sampled-row matvec parity).
  config2_x_kaXX.npz 20 480-element icosphere(5), ka in {0.25, 2, 8}: the plane-wave right-hand side is
                     recomputed by the test; the fixture holds the LAPACK (LU) solution of the oracle's
                     matrix and the oracle GMRES(50, 1e-10) iteration / restart counts.

    python tests/golden/make_golden_large.py [config4] [config2]

config2 takes ~8 minutes per frequency on 8 cores (assembly 92 s, LU 4-5 min, GMRES 25 s).
"""
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from math_audio_b200.incident import IncidentField  # noqa: E402
from math_audio_b200.mesh import generate_geodesic_sphere_mesh, generate_icosphere_mesh  # noqa: E402
from math_audio_b200.types import PhysicsParams  # noqa: E402
from oracle import oracle as orc  # noqa: E402

OUT = Path(__file__).resolve().parent

CONFIG4_NRANKS = 8
CONFIG4_STRIDE = 32
CONFIG4_NEAREST = 256


def config4_sample_rows(n: int):
    chunk = (n + CONFIG4_NRANKS - 1) // CONFIG4_NRANKS
    rows = []
    for p in range(CONFIG4_NRANKS):
        b, e = p * chunk, min(n, (p + 1) * chunk)
        rows += [b, b + (e - b) // 3, b + 2 * (e - b) // 3 + 1, e - 1]
    return np.array(rows, dtype=np.int64)


def config4_probe_vector(n: int) -> np.ndarray:
    rng = np.random.default_rng(1234)
    return rng.standard_normal(n) + 1j * rng.standard_normal(n)


def make_config4():
    a, ka = 1.0, 16.0
    mesh = generate_geodesic_sphere_mesh(a, 78)
    n = mesh.num_dofs
    ph = PhysicsParams.from_wave_number(ka / a)
    beta, _ = ph.burton_miller_beta_adaptive(a)
    rows = config4_sample_rows(n)
    xprobe = config4_probe_vector(n)
    cols = np.zeros((len(rows), CONFIG4_NEAREST + (n + CONFIG4_STRIDE - 1) // CONFIG4_STRIDE), dtype=np.int64)
    vals = np.zeros(cols.shape, dtype=np.complex128)
    rowdot = np.zeros(len(rows), dtype=np.complex128)
    rhs = np.zeros(len(rows), dtype=np.complex128)
    rownorm = np.zeros(len(rows))
    t0 = time.time()
    for i, r in enumerate(rows):
        A, rh, _ = orc.assemble(mesh, ph.wave_number, beta, row_begin=int(r), row_end=int(r) + 1)
        d = np.linalg.norm(mesh.center - mesh.center[r], axis=1)
        near = np.argsort(d, kind="stable")[:CONFIG4_NEAREST]
        c = np.concatenate([near, np.arange(0, n, CONFIG4_STRIDE)])
        cols[i] = c
        vals[i] = A[0, c]
        rowdot[i] = np.dot(A[0], xprobe)
        rhs[i] = rh[0]
        rownorm[i] = np.linalg.norm(A[0])
    np.savez_compressed(OUT / "config4_rows.npz", rows=rows, cols=cols, vals=vals, rowdot=rowdot, rownorm=rownorm, rhs=rhs, k=ph.wave_number,
                        beta=beta, nu=78, a=a, n=n)
    print(f"config4_rows: {len(rows)} rows x {cols.shape[1]} columns in {time.time() - t0:.1f} s")


def make_config2(kas=(0.25, 2.0, 8.0)):
    import scipy.linalg as sla

    a = 0.1
    mesh = generate_icosphere_mesh(a, 5)
    inc = IncidentField.plane_wave_z()
    for ka in kas:
        t0 = time.time()
        ph = PhysicsParams.from_wave_number(ka / a)
        beta, _ = ph.burton_miller_beta_adaptive(a)
        A, rhs0, _ = orc.assemble(mesh, ph.wave_number, beta)
        b = rhs0 + inc.compute_rhs_with_beta(mesh.center, mesh.normal, ph, beta)
        t1 = time.time()
        xg, info = orc.gmres(A, b, max_iterations=1000, restart=50, tolerance=1e-10)
        t2 = time.time()
        lu, piv = sla.lu_factor(A, overwrite_a=True, check_finite=False)
        x = sla.lu_solve((lu, piv), b, check_finite=False)
        t3 = time.time()
        dx = float(np.linalg.norm(xg - x) / np.linalg.norm(x))
        tag = f"{ka:g}".replace(".", "p")
        np.savez_compressed(OUT / f"config2_x_ka{tag}.npz", x=x, ka=ka, a=a, sub=5, k=ph.wave_number, beta=beta,
                            iterations=info["iterations"], restarts=info["restarts"], residual=info["residual"],
                            gmres_vs_lu=dx)
        print(f"config2 ka={ka}: assemble {t1 - t0:.0f} s, gmres {t2 - t1:.0f} s ({info}), LU {t3 - t2:.0f} s, |x_gmres - x_lu|/|x_lu| = {dx:.2e}")


if __name__ == "__main__":
    what = sys.argv[1:] or ["config4", "config2"]
    if "config4" in what:
        make_config4()
    if "config2" in what:
        make_config2()
