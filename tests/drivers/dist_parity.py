"""Row-sharded parity check, one process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tests/drivers/dist_parity.py

Every rank assembles its row block, the ranks solve together (ZGEMV epilogue over peer memory or
NCCL all-gather, BEMB200_PEER_FUSED=0 forces the latter) and each rank compares against the CPU
oracle: slab entries (rel 1e-10), operator apply, GMRES iteration count (identical), solution
(rel 1e-8) and bit-identity of the solution across ranks (replicated Arnoldi).  Exit code != 0 on
any violation.  The oracle is the checker here, never the thing measured.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import torch.distributed as dist

from math_audio_b200 import bem, dist as bdist
from math_audio_b200.incident import IncidentField
from math_audio_b200.mesh import generate_icosphere_mesh
from math_audio_b200.types import PhysicsParams
from oracle import oracle as orc

rank, local, world = bdist.env_rank()
torch.cuda.set_device(local)
bdist.init_process_group("nccl")
dev = torch.device("cuda", local)
nid = bdist.broadcast_bytes(bem.Context.nccl_unique_id() if rank == 0 else None, 128, 0, device=dev)
ctx = bem.Context(local, rank, world, nid)
a = 0.1
failures = []
for sub, ka in [(3, 1.0), (3, 8.0), (4, 2.0)]:
    mesh = generate_icosphere_mesh(a, sub)
    if sub == 3:
        mesh.is_eval[-7:] = 1  # n = 1273: uneven split
    n = mesh.num_dofs
    ph = PhysicsParams.from_wave_number(ka / a)
    beta, _ = ph.burton_miller_beta_adaptive(a)
    system = bem.build_tbem_system_with_beta(mesh, ph, beta, ctx=ctx)
    r0, r1 = system.matrix.local_rows
    Ao, rhso, _ = orc.assemble(mesh, ph.wave_number, beta)
    Aloc = system.matrix.rows()
    err = float(np.max(np.abs(Aloc - Ao[r0:r1]) / np.abs(Ao[r0:r1]))) if r1 > r0 else 0.0
    b = system.rhs_full() + IncidentField.plane_wave_z().compute_rhs_with_beta(mesh.center[:n], mesh.normal[:n], ph, beta)
    op = bem.DenseOperator(system)
    x = np.random.default_rng(1).standard_normal(n) + 1j * np.random.default_rng(2).standard_normal(n)
    y = op.apply(x)
    apply_err = float(np.linalg.norm(y - Ao @ x) / np.linalg.norm(Ao @ x))
    sol = bem.gmres(op, b, bem.GmresConfig(1000, 50, 1e-10))
    xo, io = orc.gmres(Ao, b, max_iterations=1000, restart=50, tolerance=1e-10)
    dx = float(np.linalg.norm(sol.x - xo) / np.linalg.norm(xo))
    xs = torch.from_numpy(sol.x.view(np.float64).copy()).to(dev)
    gl = [torch.zeros_like(xs) for _ in range(world)]
    dist.all_gather(gl, xs)
    same = all(bool((g == gl[0]).all()) for g in gl)
    print(f"[rank {rank}] sub={sub} ka={ka} rows=[{r0},{r1}) entry_err={err:.2e} apply_err={apply_err:.2e} "
          f"it={sol.iterations}/{io['iterations']} dx={dx:.2e} ranks_bit_identical={same} peer={ctx.peer_exchange_active()}", flush=True)
    # the other solver of BemSolver::solve_dense_system on the same sharded operator
    sb = bem.bicgstab(op, b, bem.BiCgstabConfig(1000, 1e-10, 0))
    xb, ib = orc.bicgstab(Ao, b, max_iterations=1000, tolerance=1e-10)
    dxb = float(np.linalg.norm(sb.x - xb) / np.linalg.norm(xb))
    print(f"[rank {rank}]   bicgstab it={sb.iterations}/{ib['iterations']} converged={sb.converged} dx={dxb:.2e}", flush=True)
    # BiCGSTAB's iteration count is rounding sensitive on ill-conditioned systems (tree vs sequential inner products):
    # both sides must converge to the same solution, the counts only have to be of the same order
    if not sb.converged or not ib["converged"] or abs(sb.iterations - ib["iterations"]) > 0.25 * ib["iterations"] + 2 or dxb > 1e-7:
        failures.append(f"bicgstab {sb.iterations} vs {ib['iterations']} dx {dxb}")
    # block-Jacobi (additive Schwarz, overlap 0) on rank-aligned diagonal blocks: every rank inverts and applies its own blocks
    from oracle import schwarz_oracle as so
    parts = bem.schwarz_partition_aligned(n, world, 96)
    pre = bem.AdditiveSchwarzPreconditioner.from_operator(op, subdomains=parts)
    dsw = so.DenseSchwarz(Ao, subdomains=[q.astype(np.int64) for q in parts])
    zerr = float(np.linalg.norm(pre.apply(x) - dsw.apply(x)) / np.linalg.norm(dsw.apply(x)))
    sp = bem.gmres_preconditioned(op, pre, b, bem.GmresConfig(1000, 50, 1e-10))
    xp, ip = orc.gmres_preconditioned_cb(lambda v: Ao @ v, dsw.apply, n, b, max_iterations=1000, restart=50, tolerance=1e-10)
    dxp = float(np.linalg.norm(sp.x - xp) / np.linalg.norm(xp))
    print(f"[rank {rank}]   block-jacobi {len(parts)} blocks: apply_err={zerr:.2e} it={sp.iterations}/{ip['iterations']} "
          f"(plain {sol.iterations}) dx={dxp:.2e}", flush=True)
    if zerr > 1e-11 or not sp.converged or sp.iterations != ip["iterations"] or dxp > 1e-8:
        failures.append(f"block-jacobi apply {zerr} it {sp.iterations} vs {ip['iterations']} dx {dxp}")
    pre.close()
    # several right-hand sides in lockstep on the row-sharded operator (block matvec per rank + all-gather of the block)
    Bm = np.stack([b, b[::-1].copy(), 1j * b])
    sb3, _st = bem.gmres_batched(op, Bm, bem.GmresConfig(1000, 50, 1e-10))
    Yb, _ms = bem.apply_block(op, Bm)
    berr = float(np.linalg.norm(Yb - Bm @ Ao.T) / np.linalg.norm(Bm @ Ao.T))
    for i3, s3 in enumerate(sb3):
        x3, i3o = orc.gmres(Ao, Bm[i3], max_iterations=1000, restart=50, tolerance=1e-10)
        d3 = float(np.linalg.norm(s3.x - x3) / np.linalg.norm(x3))
        if not s3.converged or s3.iterations != i3o["iterations"] or d3 > 1e-8:
            failures.append(f"batched rhs {i3}: it {s3.iterations} vs {i3o['iterations']} dx {d3}")
    print(f"[rank {rank}]   batched 3 rhs: it={[s3.iterations for s3 in sb3]} apply_block err={berr:.2e}", flush=True)
    if berr > 1e-12:
        failures.append(f"apply_block {berr}")
    if err > 1e-10:
        failures.append(f"entries {err}")
    if apply_err > 1e-12:
        failures.append(f"apply {apply_err}")
    if sol.iterations != io["iterations"] or not sol.converged:
        failures.append(f"iterations {sol.iterations} vs {io['iterations']}")
    if dx > 1e-8:
        failures.append(f"solution {dx}")
    if not same:
        failures.append("ranks disagree bitwise")
# room-acoustics matrix, row-sharded: slab parity + a sharded GMRES solve (reference settings) against the oracle
from math_audio_b200 import room
from oracle import room_oracle as ro

rmesh = room.RectangularRoom(3.1, 2.3, 2.0).generate_mesh(4)
st = room.StagedRoomMesh(rmesh, ctx)
kk = room.wavenumber(125.0, 343.0)
rc, rn, ra = ro.element_data(rmesh.nodes, rmesh.elements)
Ar = ro.build_bem_matrix(rc, rn, ra, kk)
mat = room.build_bem_matrix_parallel(st, kk)
r0, r1 = mat.local_rows
slab = mat.rows()
rerr = float(np.max(np.abs(slab - Ar[r0:r1]) / np.max(np.abs(Ar[r0:r1]), axis=1, keepdims=True)))
srcs = [room.Source.omnidirectional([1.0, 0.8, 1.1], 1.0)]
xr = room.solve_bem_system(st, srcs, kk, 125.0, reuse=mat)
br = ro.incident_field_derivative(rc, rn, [dict(position=[1.0, 0.8, 1.1], amplitude=1.0, directivity=None, crossover=dict(kind="fullrange"))], kk, 125.0)
xro, iro = orc.gmres(Ar, br, max_iterations=100, restart=50, tolerance=1e-6)
dxr = float(np.linalg.norm(xr - xro) / np.linalg.norm(xro))
print(f"[rank {rank}] room n={st.n} rows=[{r0},{r1}) rownorm_err={rerr:.2e} gmres dx={dxr:.2e}", flush=True)
if rerr > 1e-12 or dxr > 1e-8:
    failures.append(f"room {rerr} {dxr}")
if os.environ.get("BEMB200_EXPECT_PEER") == "1" and not ctx.peer_exchange_active():
    failures.append("peer-memory exchange not active")
dist.barrier()
dist.destroy_process_group()
if failures:
    print(f"[rank {rank}] FAIL: {failures}", flush=True)
    sys.exit(1)
print(f"[rank {rank}] OK", flush=True)
