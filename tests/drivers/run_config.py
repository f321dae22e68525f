#!/usr/bin/env python3
"""Run one of the BASELINE.json configs 3-4 on N GPUs (torchrun, one process per GPU), check
parity by sampled rows against the CPU oracle and by an independent residual, and write a
JSON summary to gpurun_out/config<k>_g<N>.json (copied to profiles/ by hand).

    torchrun --nproc-per-node N tests/drivers/run_config.py --config 4 [--rows-checked 64]

config 3: "cabinet" 0.32 x 0.44 x 0.64 m closed Quad4 box, 64x88x128 -> 50 176 elements, piston
          (full-length velocity BC v=1 within 80 mm of the front-wall centre), f = 1 kHz, beta = i/k.
config 4: rigid geodesic sphere nu = 78 -> 121 680 Tri3, a = 1 m, ka = 16, adaptive beta (16 i/k).
"""
import argparse
import json
import math
import os
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))


def build_case(cfg, scale):
    from math_audio_b200.mesh import generate_box_mesh_quad, generate_geodesic_sphere_mesh
    from math_audio_b200.types import PhysicsParams

    if cfg == 3:
        nx, ny, nz = int(64 * scale), int(88 * scale), int(128 * scale)
        mesh = generate_box_mesh_quad(0.32, 0.44, 0.64, nx, ny, nz)
        front = (np.abs(mesh.center[:, 1] + 0.22) < 1e-9) & (np.hypot(mesh.center[:, 0], mesh.center[:, 2]) < 0.08)
        v = np.zeros((mesh.n_elem, 4), dtype=np.complex128)
        v[front] = 1.0
        mesh.set_velocity_bc(v)
        mesh.bc_len[~front] = 1
        ph = PhysicsParams.new(1000.0, 343.0, 1.21, False)
        beta = ph.burton_miller_beta()
        return mesh, ph, beta, dict(name="cabinet_quad4", piston_elements=int(front.sum()), grid=[nx, ny, nz], nq=16), None
    if cfg == 4:
        nu = max(2, int(round(78 * scale)))
        a = 1.0
        mesh = generate_geodesic_sphere_mesh(a, nu)
        ph = PhysicsParams.from_wave_number(16.0 * scale / a)
        beta, bscale = ph.burton_miller_beta_adaptive(a)
        return mesh, ph, beta, dict(name="geodesic_sphere", nu=nu, ka=16.0 * scale, beta_scale=bscale, nq=13), "plane_wave_z"
    raise SystemExit("config must be 3 or 4")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", type=int, required=True)
    ap.add_argument("--scale", type=float, default=1.0, help="linear mesh scale (1.0 = the config's size)")
    ap.add_argument("--rows-checked", type=int, default=64)
    ap.add_argument("--tol", type=float, default=1e-10)
    ap.add_argument("--reps", type=int, default=2)
    args = ap.parse_args()

    import torch
    import torch.distributed as dist

    from math_audio_b200 import bem
    from math_audio_b200 import dist as bdist
    from math_audio_b200.incident import IncidentField
    from oracle import oracle as orc

    if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
        os.environ.pop("NCCL_DEBUG")  # keep NCCL's version banner off stdout (ONE JSON line)
    rank, local, world = bdist.env_rank()
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    nid = None
    if world > 1:
        bdist.init_process_group("nccl")
        nid = bdist.broadcast_bytes(bem.Context.nccl_unique_id() if rank == 0 else None, 128, 0, device=dev)
    ctx = bem.Context(local, rank, world, nid)

    t0 = time.perf_counter()
    mesh, ph, beta, meta, incident = build_case(args.config, args.scale)
    n = mesh.num_dofs
    t_mesh = time.perf_counter() - t0
    r0, r1 = ctx.partition(n)
    staged = bem.StagedMesh(mesh, ctx)
    system = None
    asm = []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        system = bem.build_tbem_system_with_beta(staged, ph, beta, reuse=system, fetch_rhs=False)
        torch.cuda.synchronize()
        asm.append(dict(wall_s=time.perf_counter() - t0, **system.matrix.assembly_stats()))
    rhs = system.rhs_full()
    if incident == "plane_wave_z":
        b = rhs + IncidentField.plane_wave_z().compute_rhs_with_beta(mesh.center, mesh.normal, ph, beta)
    else:
        b = rhs
    op = bem.DenseOperator(system)
    cfg = bem.GmresConfig(max_iterations=1000, restart=50, tolerance=args.tol)
    sols = []
    for _ in range(args.reps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        sol = bem.gmres(op, b, cfg)
        torch.cuda.synchronize()
        sols.append(dict(wall_s=time.perf_counter() - t0, iterations=sol.iterations, restarts=sol.restarts,
                         residual=sol.residual, converged=sol.converged, **system.matrix.solver_stats()))
    # independent residual through the operator boundary
    res = float(np.linalg.norm(b - op.apply(sol.x)) / np.linalg.norm(b))
    # sampled-row entry parity (+ rhs parity) against the oracle, rows of THIS rank
    rng = np.random.default_rng(1000 + rank)
    per_rank = max(2, args.rows_checked // world)
    rows = sorted(set([r0, r1 - 1] + [int(x) for x in rng.integers(r0, r1, per_rank - 2)]))
    worst_rel = worst_rown = worst_rhs = 0.0
    t0 = time.perf_counter()
    xs = np.random.default_rng(1234).standard_normal(n) + 1j * np.random.default_rng(4321).standard_normal(n)
    ys = op.apply(xs)
    worst_mv = 0.0
    for r in rows:
        Ar = system.matrix.rows(r, r + 1)
        Ao, rhso, _ = orc.assemble(mesh, ph.wave_number, beta, row_begin=r, row_end=r + 1)
        sc = np.abs(Ao).max()
        worst_rown = max(worst_rown, float(np.abs(Ar - Ao).max() / sc))
        big = np.abs(Ao) > 1e-9 * sc
        worst_rel = max(worst_rel, float((np.abs(Ar - Ao)[big] / np.abs(Ao)[big]).max()))
        if np.abs(rhso).max() > 0:
            worst_rhs = max(worst_rhs, float(abs(rhs[r] - rhso[0]) / abs(rhso[0])))
        worst_mv = max(worst_mv, float(abs(ys[r] - (Ao @ xs)[0]) / (np.linalg.norm(Ao) * np.linalg.norm(xs))))
    t_oracle = time.perf_counter() - t0
    errs = torch.tensor([worst_rel, worst_rown, worst_rhs, worst_mv], dtype=torch.float64, device=dev)
    tim = torch.tensor([asm[-1]["total_ms"], asm[-1]["far_ms"], sols[-1]["wall_s"], sols[-1]["matvec_ms"], asm[-1]["wall_s"]],
                       dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(errs, op=dist.ReduceOp.MAX)
        dist.all_reduce(tim, op=dist.ReduceOp.MAX)
    if rank == 0:
        nloc = r1 - r0
        far_flop = (68.0 * meta["nq"] + 40.0) * nloc * (n - 1)
        mv_bytes = 16.0 * nloc * n + 16.0 * n + 16.0 * nloc
        out = dict(config=args.config, meta=meta, n_elements=int(n), n_gpus=world, rows_per_gpu=int(nloc),
                   matrix_gb_per_gpu=16.0 * nloc * n / 1e9, mesh_build_s=t_mesh,
                   assemble=dict(max_total_ms=float(tim[0]), max_far_ms=float(tim[1]), wall_s=float(tim[4]),
                                 far_tflops_per_gpu=far_flop / (float(tim[1]) * 1e-3) / 1e12 if tim[1] > 0 else None,
                                 far_frac_of_nominal_fp64=far_flop / (float(tim[1]) * 1e-3) / 37.22496e12 if tim[1] > 0 else None,
                                 near_pairs_rank0=asm[-1]["near_pairs"], special_pairs_rank0=asm[-1]["special_pairs"]),
                   solve=dict(wall_s=float(tim[2]), iterations=sols[-1]["iterations"], restarts=sols[-1]["restarts"],
                              reported_residual=sols[-1]["residual"], converged=sols[-1]["converged"],
                              independent_residual=res, matvecs=sols[-1]["matvecs"], matvec_ms_total=float(tim[3]),
                              matvec_gbs_per_gpu=mv_bytes * sols[-1]["matvecs"] / (float(tim[3]) * 1e-3) / 1e9,
                              matvec_frac_of_measured_hbm=mv_bytes * sols[-1]["matvecs"] / (float(tim[3]) * 1e-3) / 6451.8e9),
                   seconds_per_frequency=float(tim[4]) + float(tim[2]),
                   parity=dict(rows_checked=len(rows) * world, max_rel_entry_err=float(errs[0]), max_rownorm_entry_err=float(errs[1]),
                               max_rhs_rel_err=float(errs[2]), max_matvec_row_err=float(errs[3]), oracle_seconds=t_oracle),
                   runs=dict(assemble=asm, solve=sols))
        path = ROOT / "gpurun_out" / f"config{args.config}_g{world}_s{args.scale:g}.json"
        path.parent.mkdir(exist_ok=True)
        path.write_text(json.dumps(out, indent=1))
        print(json.dumps({k: out[k] for k in ("config", "n_elements", "n_gpus", "seconds_per_frequency", "parity")}))
        print(json.dumps(out["assemble"]))
        print(json.dumps(out["solve"]))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
