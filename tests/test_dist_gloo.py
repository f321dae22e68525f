"""world_size-2 gloo test of the row-sharded design on CPU ("sharded matvec, replicated
Arnoldi", csrc/gmres.cu): every rank assembles only its row block, the matvec output is
all-gathered in the padded layout, and every rank runs the same GMRES on the full vectors.
The result must be bit-identical on all ranks and to the single-process solve."""
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent

WORKER = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.environ["REPO_ROOT"])
from math_audio_b200 import dist as bdist
from math_audio_b200.mesh import generate_icosphere_mesh
from math_audio_b200.types import PhysicsParams
from oracle import oracle as orc

rank, world = bdist.init_process_group("gloo")
assert world == 2
ident = bdist.broadcast_bytes(bytes(range(128)) if rank == 0 else None, 128)
assert ident == bytes(range(128))
mesh = generate_icosphere_mesh(0.1, 1)           # N = 80 -> odd split below via n = 77 dofs
mesh.is_eval[-3:] = 1
n = mesh.num_dofs
ph = PhysicsParams.from_wave_number(20.0)
beta, _ = ph.burton_miller_beta_adaptive(0.1)
r0, r1 = bdist.partition(n, world, rank)
chunk, npad = bdist.gather_layout(n, world)
assert (r0, r1) == ((0, 39) if rank == 0 else (39, 77)) and npad == 78
A_loc, rhs_loc, _ = orc.assemble(mesh, ph.wave_number, beta, row_begin=r0, row_end=r1)
rhs = bdist.allgather_rows(rhs_loc, n)
inc, _ = orc.incident_rhs(0, [0, 0, 1.0], 1.0, mesh.center[:n], mesh.normal[:n], ph.wave_number, beta)
b = rhs + inc
x, info = orc.gmres_op(lambda v: bdist.allgather_rows(orc.zgemv(A_loc, v, nthreads=1), n), n, b,
                       max_iterations=50, restart=10, tolerance=1e-10)
# row-sharded block-Jacobi (csrc/schwarz.cu): rank-aligned subdomains, every rank factors and applies ONLY the blocks of its own
# rows on its slab of A v, then the preconditioned slabs are gathered -- must equal the global additive Schwarz preconditioner
from math_audio_b200 import bem
from oracle import schwarz_oracle as so
parts = bem.schwarz_partition_aligned(n, world, 10)
mine = [p.astype(np.int64) for p in parts if r0 <= int(p[0]) and int(p[-1]) < r1]
assert sum(len(p) for p in mine) == r1 - r0
local = so.DenseSchwarz(A_loc[:, r0:r1], subdomains=[p - r0 for p in mine])
def precond(v):
    return bdist.allgather_rows(local.apply(v[r0:r1]), n)
xp, infop = orc.gmres_preconditioned_cb(lambda v: bdist.allgather_rows(orc.zgemv(A_loc, v, nthreads=1), n), precond, n, b,
                                        max_iterations=50, restart=10, tolerance=1e-10)
np.save(os.path.join(os.environ["OUT_DIR"], f"xp_{rank}.npy"), xp)
np.save(os.path.join(os.environ["OUT_DIR"], f"infop_{rank}.npy"), np.array([infop["iterations"], infop["restarts"], int(infop["converged"])]))
np.save(os.path.join(os.environ["OUT_DIR"], f"x_{rank}.npy"), x)
np.save(os.path.join(os.environ["OUT_DIR"], f"info_{rank}.npy"), np.array([info["iterations"], info["restarts"], int(info["converged"])]))
if rank == 0:
    A, rhs_full, _ = orc.assemble(mesh, ph.wave_number, beta)
    xs, infos = orc.gmres(A, rhs_full + inc, max_iterations=50, restart=10, tolerance=1e-10, nthreads=1)
    np.save(os.path.join(os.environ["OUT_DIR"], "x_single.npy"), xs)
    np.save(os.path.join(os.environ["OUT_DIR"], "info_single.npy"), np.array([infos["iterations"], infos["restarts"], int(infos["converged"])]))
    glob = so.DenseSchwarz(A, subdomains=[p.astype(np.int64) for p in parts])
    xps, infops = orc.gmres_preconditioned_cb(lambda v: orc.zgemv(A, v, nthreads=1), glob.apply, n, rhs_full + inc,
                                              max_iterations=50, restart=10, tolerance=1e-10)
    np.save(os.path.join(os.environ["OUT_DIR"], "xp_single.npy"), xps)
    np.save(os.path.join(os.environ["OUT_DIR"], "infop_single.npy"), np.array([infops["iterations"], infops["restarts"], int(infops["converged"])]))
'''


def test_sharded_matvec_replicated_arnoldi_gloo(tmp_path, orc):
    worker = tmp_path / "worker.py"
    worker.write_text(WORKER)
    env = dict(os.environ, REPO_ROOT=str(ROOT), OUT_DIR=str(tmp_path), MASTER_ADDR="127.0.0.1", MASTER_PORT="29533",
               WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(worker)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r))) for r in range(2)]
    for p in procs:
        assert p.wait(timeout=300) == 0
    x0, x1, xs = (np.load(tmp_path / f) for f in ("x_0.npy", "x_1.npy", "x_single.npy"))
    i0, i1, isg = (np.load(tmp_path / f) for f in ("info_0.npy", "info_1.npy", "info_single.npy"))
    assert (x0 == x1).all(), "ranks diverged"
    assert (x0 == xs).all(), "sharded solve differs from the single-process solve"
    assert (i0 == i1).all() and (i0 == isg).all() and i0[2] == 1 and i0[1] >= 1  # restarted at least once
    p0, p1, ps = (np.load(tmp_path / f) for f in ("xp_0.npy", "xp_1.npy", "xp_single.npy"))
    j0, j1, js = (np.load(tmp_path / f) for f in ("infop_0.npy", "infop_1.npy", "infop_single.npy"))
    assert (p0 == p1).all() and (p0 == ps).all(), "row-sharded block-Jacobi differs from the global preconditioner"
    assert (j0 == j1).all() and (j0 == js).all() and j0[2] == 1 and j0[0] < i0[0]  # and it does precondition


RIDE_ALONG_WORKER = r'''
import json, os, sys
sys.path.insert(0, os.environ["REPO_ROOT"])
import torch
from math_audio_b200 import dist as bdist
import bench

rank, world = bdist.init_process_group("gloo")
dev = torch.device("cpu")
def boom():
    raise RuntimeError("golden file mismatch")
res = {}
# 1. nobody fails: the block comes back untouched
res["ok"] = bench.ride_along("blk", lambda: {"v": rank}, world, dev)
# 2. every rank fails the same way: every rank carries the error and goes on to print its line
res["all"] = bench.ride_along("blk", boom, world, dev)
# 3. rank 1 fails alone but rank 0 still arrives: rank 0 learns about it, its own block is kept beside the error
res["one"] = bench.ride_along("blk", (boom if rank == 1 else (lambda: {"v": 0})), world, dev)
json.dump(res, open(os.path.join(os.environ["OUT_DIR"], f"ride_{rank}.json"), "w"))
# 4. rank 1 fails alone and rank 0 never arrives (it sits in a collective of the block): rank 1 ends the job after wait_s
if rank == 1:
    bench.ride_along("blk", boom, world, dev, wait_s=1.0)
    sys.exit(0)  # not reached
else:
    import time
    time.sleep(20)
'''


def test_bench_side_blocks_cannot_take_the_headline_line_with_them(tmp_path):
    """bench.ride_along on 2 gloo ranks: errors become {"error": ...} on every rank (agreed with one all-reduce); a rank that
    failed alone and is not joined within its time limit exits non-zero instead of hanging the job."""
    worker = tmp_path / "ride.py"
    worker.write_text(RIDE_ALONG_WORKER)
    env = dict(os.environ, REPO_ROOT=str(ROOT), OUT_DIR=str(tmp_path), MASTER_ADDR="127.0.0.1", MASTER_PORT="29537",
               WORLD_SIZE="2", OMP_NUM_THREADS="1")
    procs = [subprocess.Popen([sys.executable, str(worker)], env=dict(env, RANK=str(r), LOCAL_RANK=str(r)), stderr=subprocess.PIPE, text=True)
             for r in range(2)]
    err1 = procs[1].communicate(timeout=120)[1]
    assert procs[1].returncode == 1 and "failed on this rank only" in err1
    procs[0].kill()
    procs[0].communicate()
    import json

    r0, r1 = (json.load(open(tmp_path / f"ride_{r}.json")) for r in range(2))
    assert r0["ok"] == {"v": 0} and r1["ok"] == {"v": 1}
    assert r0["all"] == r1["all"] == {"error": "RuntimeError: golden file mismatch"}
    assert r1["one"] == {"error": "RuntimeError: golden file mismatch"}
    assert r0["one"] == {"error": "blk failed on another rank", "this_rank": {"v": 0}}
