"""Block-Jacobi (additive Schwarz, overlap 0) on one GPU: GMRES iterations and time with and without the preconditioner for
several block sizes, contiguous blocks vs spatial clusters.  usage: precond_probe.py [sub=5] [ka ...]"""
import sys
import time

sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
import numpy as np

from math_audio_b200 import bem
from math_audio_b200.incident import IncidentField
from math_audio_b200.mesh import generate_geodesic_sphere_mesh, generate_icosphere_mesh
from math_audio_b200.types import PhysicsParams

sub = int(sys.argv[1]) if len(sys.argv) > 1 else 5
kas = [float(v) for v in sys.argv[2:]] or [2.0, 8.0]
import os
SIZES = [int(v) for v in os.environ.get("PROBE_SIZES", "32,64,128,256,512,1024").split(",")]
KINDS = os.environ.get("PROBE_KINDS", "contiguous,spatial").split(",")
a = 0.1
mesh = generate_icosphere_mesh(a, sub) if sub < 10 else generate_geodesic_sphere_mesh(a, sub)
st = bem.StagedMesh(mesh)
n = st.num_dofs
cfg = bem.GmresConfig(1000, 50, 1e-10)
for ka in kas:
    ph = PhysicsParams.from_wave_number(ka / a)
    beta, _ = ph.burton_miller_beta_adaptive(a)
    system = bem.build_tbem_system_with_beta(st, ph, beta)
    b = system.rhs_full() + IncidentField.plane_wave_z().compute_rhs_with_beta(mesh.center[:n], mesh.normal[:n], ph, beta)
    op = bem.DenseOperator(system)
    bem.gmres(op, b, bem.GmresConfig(1, 2, 1e-10))
    t0 = time.perf_counter(); plain = bem.gmres(op, b, cfg); t_plain = time.perf_counter() - t0
    print(f"n={n} ka={ka}: plain GMRES {plain.iterations} it in {t_plain*1e3:.1f} ms", flush=True)
    jac = bem.DiagonalPreconditioner.from_operator(op)
    t0 = time.perf_counter(); sj = bem.gmres_preconditioned(op, jac, b, cfg); tj = time.perf_counter() - t0
    print(f"   jacobi: {sj.iterations} it in {tj*1e3:.1f} ms", flush=True)
    for bs in SIZES:
        for kind in KINDS:
            parts = (bem.schwarz_partition_aligned(n, 1, bs) if kind == "contiguous" else
                     bem.spatial_subdomains(mesh.center[:n], 1, bs) if kind == "spatial" else
                     bem.voronoi_subdomains(mesh.center[:n], 1, bs))
            t0 = time.perf_counter()
            pre = bem.AdditiveSchwarzPreconditioner.from_operator(op, subdomains=parts)
            t_build = time.perf_counter() - t0
            stt = pre.stats()
            t0 = time.perf_counter(); sol = bem.gmres_preconditioned(op, pre, b, cfg); t_sol = time.perf_counter() - t0
            res = np.linalg.norm(op.apply(sol.x) - b) / np.linalg.norm(b)
            dx = np.linalg.norm(sol.x - plain.x) / np.linalg.norm(plain.x)
            print(f"   block-jacobi {kind:10s} size<={bs:4d} ({stt['num_subdomains']} blocks, {stt['inverse_bytes']/1e6:.0f} MB, set-up {stt['factor_ms']:.1f} ms "
                  f"device / {t_build*1e3:.1f} ms wall): {sol.iterations} it in {t_sol*1e3:.1f} ms, total {1e3*(t_build+t_sol):.1f} ms; "
                  f"true residual {res:.2e}, dx vs plain {dx:.1e}", flush=True)
            pre.close()
    system.matrix.close()
