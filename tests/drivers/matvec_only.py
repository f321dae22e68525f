"""One foreground assembly + a few operator applies at 20 480 elements (for an ncu capture of zgemv_kernel)."""
import sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
import numpy as np
from math_audio_b200 import bem
from math_audio_b200.mesh import generate_icosphere_mesh
from math_audio_b200.types import PhysicsParams
a = 0.1
mesh = generate_icosphere_mesh(a, 5)
st = bem.StagedMesh(mesh)
ph = PhysicsParams.from_wave_number(2.0 / a)
beta, _ = ph.burton_miller_beta_adaptive(a)
sysg = bem.build_tbem_system_with_beta(st, ph, beta, fetch_rhs=False)
op = bem.DenseOperator(sysg)
x = np.random.default_rng(1).standard_normal(mesh.num_dofs) + 0j
for rep in range(5):
    y = op.apply(x)
    print(sysg.matrix.solver_stats())
