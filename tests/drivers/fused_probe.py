"""Probe of the persistent fused GMRES kernel: parity against the oracle on small meshes, then timing at
20 480 elements against the per-iteration kernels (BEMB200_GMRES_FUSED=0 in a child process).

    python tests/drivers/fused_probe.py [--big] [--legacy]
"""
import ctypes as C
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))

from math_audio_b200 import _capi, bem  # noqa: E402
from math_audio_b200.incident import IncidentField  # noqa: E402
from math_audio_b200.mesh import generate_icosphere_mesh  # noqa: E402
from math_audio_b200.types import PhysicsParams  # noqa: E402


def fused_times():
    lib = _capi.lib()
    f = lib.bemb200_debug_fused_times
    f.restype = None
    a, b, c, d = C.c_double(), C.c_double(), C.c_double(), C.c_ulonglong()
    f(C.byref(a), C.byref(b), C.byref(c), C.byref(d))
    return a.value, b.value, c.value, d.value


def small():
    from oracle import oracle as orc

    a = 0.1
    for sub, ka, restart, tol in ((1, 0.5, 50, 1e-10), (2, 1.0, 50, 1e-10), (2, 6.0, 5, 1e-10), (3, 3.0, 50, 1e-10), (3, 8.0, 10, 1e-8),
                                  (2, 1.0, 1, 1e-3), (3, 1.0, 63, 1e-12)):
        mesh = generate_icosphere_mesh(a, sub)
        ph = PhysicsParams.from_wave_number(ka / a)
        beta, _ = ph.burton_miller_beta_adaptive(a)
        system = bem.build_tbem_system_with_beta(mesh, ph, beta)
        Ao, _, _ = orc.assemble(mesh, ph.wave_number, beta)
        b = system.rhs + IncidentField.plane_wave_z().compute_rhs_with_beta(mesh.center, mesh.normal, ph, beta)
        cfg = bem.GmresConfig(max_iterations=40, restart=restart, tolerance=tol)
        t0 = time.perf_counter()
        sol = bem.gmres(bem.DenseOperator(system), b, cfg)
        dt = time.perf_counter() - t0
        xo, io = orc.gmres(Ao, b, max_iterations=40, restart=restart, tolerance=tol)
        dx = float(np.linalg.norm(sol.x - xo) / np.linalg.norm(xo))
        ok = (sol.iterations == io["iterations"] and sol.restarts == io["restarts"] and sol.converged == io["converged"])
        print(f"n={mesh.n_elem} ka={ka} restart={restart}: it {sol.iterations}/{io['iterations']} restarts {sol.restarts}/{io['restarts']} "
              f"conv {sol.converged}/{io['converged']} res {sol.residual:.3e}/{io['residual']:.3e} dx {dx:.2e} {'OK' if ok and dx < 1e-7 else 'MISMATCH'} "
              f"({dt * 1e3:.1f} ms) fused={fused_times()}", flush=True)
        # zero right-hand side and exact initial guess
        z = bem.gmres(bem.DenseOperator(system), np.zeros_like(b), cfg)
        g = bem.gmres_with_guess(bem.DenseOperator(system), b, sol.x, cfg)
        go_x, go = orc.gmres(Ao, b, x0=sol.x, max_iterations=40, restart=restart, tolerance=tol)
        print(f"   zero rhs: it {z.iterations} conv {z.converged} res {z.residual}; warm start: it {g.iterations}/{go['iterations']} conv {g.converged}/{go['converged']}", flush=True)


def big(steps=6):
    a = 0.1
    mesh = generate_icosphere_mesh(a, 5)
    st = bem.StagedMesh(mesh)
    inc = IncidentField.plane_wave_z()
    cfg = bem.GmresConfig(max_iterations=1000, restart=50, tolerance=1e-10)
    system = None
    for i, ka in enumerate((2.0, 0.25, 8.0, 2.0, 2.0, 4.0)[:steps]):
        ph = PhysicsParams.from_wave_number(ka / a)
        beta, _ = ph.burton_miller_beta_adaptive(a)
        system = bem.build_tbem_system_with_beta(st, ph, beta, reuse=system, fetch_rhs=False)
        b = inc.compute_rhs_with_beta(mesh.center, mesh.normal, ph, beta)
        op = bem.DenseOperator(system)
        t0 = time.perf_counter()
        sol = bem.gmres(op, b, cfg)
        dt = time.perf_counter() - t0
        res = float(np.linalg.norm(b - op.apply(sol.x)) / np.linalg.norm(b))
        stt = system.matrix.solver_stats()
        print(f"ka={ka}: it {sol.iterations} restarts {sol.restarts} conv {sol.converged} res {sol.residual:.3e} true {res:.3e} wall {dt * 1e3:.2f} ms "
              f"stats {stt} fused(total, matvec, round ms, rounds)={fused_times()}", flush=True)


if __name__ == "__main__":
    if "--big" in sys.argv:
        big()
    else:
        small()
