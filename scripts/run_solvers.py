#!/usr/bin/env python3
"""The dense solvers behind BemSolver::solve_dense_system / solve_gmres / solve_cgs on one assembled
system (icosphere(5), 20 480 elements, ka = 2, beta = 4 i/k as BemSolver's default beta_scale):

    GMRES(50) tol 1e-10 (gmres.rs)  |  BiCGSTAB tol 1e-10 (bicgstab.rs)  |  CGS tol 1e-10 (cgs.rs)  |  LU (lu.rs -> cuSOLVER zgetrf/zgetrs)

Reports device/wall times, matvec counts, residuals and mutual agreement of the solutions.
Writes gpurun_out/solvers_20480.json.
"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch

    from math_audio_b200 import bem
    from math_audio_b200.incident import IncidentField
    from math_audio_b200.mesh import generate_icosphere_mesh
    from math_audio_b200.types import PhysicsParams

    sub = int(sys.argv[1]) if len(sys.argv) > 1 else 5
    a = 0.1
    mesh = generate_icosphere_mesh(a, sub)
    ph = PhysicsParams.from_wave_number(2.0 / a)
    beta = ph.burton_miller_beta_scaled(4.0)
    system = bem.build_tbem_system_with_beta(mesh, ph, beta)
    b = system.rhs_full() + IncidentField.plane_wave_z().compute_rhs_with_beta(mesh.center, mesh.normal, ph, beta)
    op = bem.DenseOperator(system)
    n = mesh.num_dofs
    out = dict(n_elements=n, ka=2.0)

    def timed(fn):
        fn()  # warm-up (workspaces, cuSOLVER handle)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        r = fn()
        torch.cuda.synchronize()
        return r, time.perf_counter() - t0

    g, tg = timed(lambda: bem.gmres(op, b, bem.GmresConfig(1000, 50, 1e-10)))
    sg = system.matrix.solver_stats()
    out["gmres"] = dict(seconds=tg, iterations=g.iterations, matvecs=sg["matvecs"], residual=g.residual, converged=g.converged)
    bc, tb = timed(lambda: bem.bicgstab(op, b, bem.BiCgstabConfig(1000, 1e-10, 0)))
    sb = system.matrix.solver_stats()
    out["bicgstab"] = dict(seconds=tb, iterations=bc.iterations, matvecs=sb["matvecs"], residual=bc.residual, converged=bc.converged,
                           matvec_gbs=(16 * n * n + 32 * n) * sb["matvecs"] / (sb["matvec_ms"] * 1e-3) / 1e9)
    cg, tc = timed(lambda: bem.cgs(op, b, bem.CgsConfig(1000, 1e-10, 0)))
    sc = system.matrix.solver_stats()
    out["cgs"] = dict(seconds=tc, iterations=cg.iterations, matvecs=sc["matvecs"], residual=cg.residual, converged=cg.converged,
                      matvec_gbs=(16 * n * n + 32 * n) * sc["matvecs"] / (sc["matvec_ms"] * 1e-3) / 1e9)
    st = {}
    x_lu, tl = timed(lambda: bem.lu_solve(system, b, stats=st))
    flops = 8.0 / 3.0 * n ** 3
    out["lu"] = dict(seconds=tl, factor_ms=st["factor_ms"], factor_tflops=flops / (st["factor_ms"] * 1e-3) / 1e12,
                     residual=float(np.linalg.norm(op.apply(x_lu) - b) / np.linalg.norm(b)))
    out["agreement"] = dict(gmres_vs_lu=float(np.linalg.norm(g.x - x_lu) / np.linalg.norm(x_lu)),
                            bicgstab_vs_lu=float(np.linalg.norm(bc.x - x_lu) / np.linalg.norm(x_lu)),
                            cgs_vs_lu=float(np.linalg.norm(cg.x - x_lu) / np.linalg.norm(x_lu)))
    (ROOT / "gpurun_out").mkdir(exist_ok=True)
    (ROOT / "gpurun_out" / f"solvers_{n}.json").write_text(json.dumps(out, indent=1))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
