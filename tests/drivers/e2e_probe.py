"""Where the host-buffer (e2e) path spends its time at 20 480 elements: per-phase wall clock of the sequential calls."""
import sys, time, numpy as np
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
import torch
from math_audio_b200 import bem
from math_audio_b200.mesh import generate_icosphere_mesh
from math_audio_b200.types import PhysicsParams
from math_audio_b200.incident import IncidentField
ctx = bem.default_context()
a = 0.1
mesh = generate_icosphere_mesh(a, 5)
n = mesh.n_elem
inc = IncidentField.plane_wave_z()
cfg = bem.GmresConfig(max_iterations=1000, restart=50, tolerance=1e-10)
sysg = None
xd = torch.zeros(n, dtype=torch.complex128, device='cuda')
for ka in (0.5, 1.0, 2.0, 2.0, 2.0, 2.0):
    ph = PhysicsParams.from_wave_number(ka / a)
    beta, _ = ph.burton_miller_beta_adaptive(a)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    st = bem.StagedMesh(mesh)
    t1 = time.perf_counter()
    sysg = bem.build_tbem_system_with_beta(st, ph, beta, reuse=sysg, fetch_rhs=False)
    t2 = time.perf_counter()
    r = sysg.rhs_full()
    t3 = time.perf_counter()
    b = r + inc.compute_rhs_with_beta(mesh.center, mesh.normal, ph, beta)
    t4 = time.perf_counter()
    op = bem.DenseOperator(sysg)
    sol = bem.gmres(op, b, cfg)
    t5 = time.perf_counter()
    ss = sysg.matrix.solver_stats()
    bd = torch.from_numpy(b).cuda()
    torch.cuda.synchronize()
    t6 = time.perf_counter()
    sold = bem.gmres_device(op, bd.data_ptr(), xd.data_ptr(), cfg)
    torch.cuda.synchronize()
    t7 = time.perf_counter()
    asm = sysg.matrix.assembly_stats()
    print(f"ka={ka} stage={1e3*(t1-t0):.2f} assemble_call={1e3*(t2-t1):.2f} (device total {asm['total_ms']:.2f}, far {asm['far_ms']:.2f}) "
          f"rhs_full={1e3*(t3-t2):.2f} host_rhs={1e3*(t4-t3):.2f} gmres_host={1e3*(t5-t4):.2f} gmres_device={1e3*(t7-t6):.2f} "
          f"it={sol.iterations} matvec_ms={ss['matvec_ms']:.2f} matvecs={ss['matvecs']}")
