"""Foreground far-kernel time at several problem sizes on one GPU (Tri3 spheres and the Quad4 cabinet box)."""
import sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
from math_audio_b200 import bem
from math_audio_b200.mesh import generate_box_mesh_quad, generate_geodesic_sphere_mesh, generate_icosphere_mesh
from math_audio_b200.types import PhysicsParams

def run(name, mesh, ph, beta, nq, rows=None):
    st = bem.StagedMesh(mesh)
    n = st.num_dofs
    rows = rows or (0, n)
    sysg = None
    best = 1e30
    for rep in range(3):
        sysg = bem.build_tbem_system_with_beta(st, ph, beta, rows=rows, reuse=sysg, fetch_rhs=False)
        s = sysg.matrix.assembly_stats()
        best = min(best, s["far_ms"])
    flop = (68.0 * nq + 40.0) * (rows[1] - rows[0]) * (n - 1)
    print(f"{name}: n={n} rows={rows[1]-rows[0]} far {best:.2f} ms = {flop / best / 1e9:.2f} TF ({flop / best / 1e9 / 37.22:.3f} of nominal) total {s['total_ms']:.2f} ms special {s['special_pairs']}", flush=True)
    sysg.matrix.close()

a = 0.1
ph = PhysicsParams.from_wave_number(2.0 / a)
beta, _ = ph.burton_miller_beta_adaptive(a)
run("icosphere(5)", generate_icosphere_mesh(a, 5), ph, beta, 13)
ph4 = PhysicsParams.from_wave_number(16.0)
b4, _ = ph4.burton_miller_beta_adaptive(1.0)
m4 = generate_geodesic_sphere_mesh(1.0, 78)
run("config4 slab (1/8 of 121680)", m4, ph4, b4, 13, rows=(0, 15210))
import numpy as np
box = generate_box_mesh_quad(0.32, 0.44, 0.64, 64, 88, 128)
front = (np.abs(box.center[:, 1] + 0.22) < 1e-9) & (np.hypot(box.center[:, 0], box.center[:, 2]) < 0.08)
v = np.zeros((box.n_elem, 4), dtype=np.complex128); v[front] = 1.0
box.set_velocity_bc(v); box.bc_len[~front] = 1
ph3 = PhysicsParams.new(1000.0, 343.0, 1.21, False)
run("config3 slab (1/2 of 50176 Quad4 + piston)", box, ph3, ph3.burton_miller_beta(), 16, rows=(0, 25088))
