#!/usr/bin/env python3
"""BASELINE config 5: icosphere(5) (20 480 elements), ka = 2, beta = 16 i/k, 32 incident plane-wave
directions (Fibonacci sphere); batched GMRES(50) tol 1e-10 on the tensor-core block matvec, compared
with 32 sequential single-RHS device solves.  Writes gpurun_out/config5.json."""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
from math_audio_b200 import bem  # noqa: E402
from math_audio_b200.incident import IncidentField  # noqa: E402
from math_audio_b200.mesh import fibonacci_directions, generate_icosphere_mesh  # noqa: E402
from math_audio_b200.types import PhysicsParams  # noqa: E402
from oracle import oracle as orc  # noqa: E402

sub = int(sys.argv[1]) if len(sys.argv) > 1 else 5
a = 0.1
mesh = generate_icosphere_mesh(a, sub)
n = mesh.n_elem
ph = PhysicsParams.from_wave_number(2.0 / a)
beta, scale = ph.burton_miller_beta_adaptive(a)
system = bem.build_tbem_system_with_beta(mesh, ph, beta)
op = bem.DenseOperator(system)
dirs = fibonacci_directions(32)
B = np.stack([system.rhs + IncidentField.plane_wave(d).compute_rhs_with_beta(mesh.center, mesh.normal, ph, beta) for d in dirs])
cfg = bem.GmresConfig(max_iterations=1000, restart=50, tolerance=1e-10)
X = np.random.default_rng(1).standard_normal((32, n)) + 1j * np.random.default_rng(2).standard_normal((32, n))
Y, kms = bem.apply_block(op, X)
y0 = op.apply(X[0])
zg = []
for _ in range(6):
    op.apply(X[0])
    zg.append(system.matrix.solver_stats()["matvec_ms"])
zgemv_ms = float(np.median(zg[1:]))
blk_err = float(np.linalg.norm(Y[0] - y0) / np.linalg.norm(y0))
bem.gmres_batched(op, B[:8], bem.GmresConfig(1, 2, 1e-10))  # warm-up
t0 = time.perf_counter()
sols, st = bem.gmres_batched(op, B, cfg)
t_batched = time.perf_counter() - t0
t0 = time.perf_counter()
singles = [bem.gmres(op, B[i], cfg) for i in range(32)]
t_seq = time.perf_counter() - t0
dx = [float(np.linalg.norm(sols[i].x - singles[i].x) / np.linalg.norm(singles[i].x)) for i in range(32)]
res = [float(np.linalg.norm(B[i] - op.apply(sols[i].x)) / np.linalg.norm(B[i])) for i in (0, 7, 31)]
# oracle rows: sampled-row check of the block matvec
rows = [0, n // 2, n - 1]
mv_err = 0.0
for r in rows:
    Ao, _, _ = orc.assemble(mesh, ph.wave_number, beta, row_begin=r, row_end=r + 1)
    mv_err = max(mv_err, float(np.abs(Y[:, r] - X @ Ao[0]).max() / (np.linalg.norm(Ao) * np.linalg.norm(X[0]))))
flops = 8.0 * n * n * 32
out = dict(config=5, n_elements=n, nrhs=32, beta_scale=scale,
           block_matvec=dict(kernel_ms=kms, tflops=flops / (kms * 1e-3) / 1e12, frac_of_nominal_fp64=flops / (kms * 1e-3) / 37.22496e12,
                             bytes_per_launch=16.0 * n * n + 32.0 * n * 32, equivalent_32_zgemv_ms=32 * zgemv_ms, zgemv_ms=zgemv_ms,
                             err_vs_zgemv=blk_err, err_vs_oracle_rows=mv_err),
           batched=dict(wall_s=t_batched, **st, iterations=[s.iterations for s in sols], restarts=[s.restarts for s in sols],
                        all_converged=all(s.converged for s in sols), max_reported_residual=max(s.residual for s in sols)),
           sequential=dict(wall_s=t_seq, iterations=[s.iterations for s in singles]),
           speedup=t_seq / t_batched, max_dx_vs_single_rhs=max(dx), independent_residuals=res,
           iterations_equal=[s.iterations for s in sols] == [s.iterations for s in singles])
(ROOT / "gpurun_out").mkdir(exist_ok=True)
(ROOT / "gpurun_out" / f"config5_sub{sub}.json").write_text(json.dumps(out, indent=1))
print(json.dumps({k: out[k] for k in ("n_elements", "block_matvec", "speedup", "max_dx_vs_single_rhs", "iterations_equal")}))
print("batched", t_batched, st, "sequential", t_seq)
