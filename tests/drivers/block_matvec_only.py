"""Block matvec alone (icosphere(5), 32 right-hand sides): kernel time of bemb200_apply_block; the target of an ncu capture."""
import sys
sys.path.insert(0, str(__import__('pathlib').Path(__file__).resolve().parents[2]))
import numpy as np
from math_audio_b200 import bem
from math_audio_b200.mesh import generate_icosphere_mesh
from math_audio_b200.types import PhysicsParams

sub = int(sys.argv[1]) if len(sys.argv) > 1 else 5
nrhs = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
a = 0.1
mesh = generate_icosphere_mesh(a, sub)
n = mesh.n_elem
ph = PhysicsParams.from_wave_number(2.0 / a)
beta, _ = ph.burton_miller_beta_adaptive(a)
system = bem.build_tbem_system_with_beta(mesh, ph, beta)
op = bem.DenseOperator(system)
X = np.random.default_rng(1).standard_normal((nrhs, n)) + 1j * np.random.default_rng(2).standard_normal((nrhs, n))
ms = []
for _ in range(reps):
    Y, k = bem.apply_block(op, X)
    ms.append(k)
y0 = op.apply(X[0])
S = ((nrhs + 7) // 8) * 8
flops = 8.0 * n * n * S
best = min(ms)
print(f"n={n} nrhs={nrhs} (S={S}): block matvec {best:.3f} ms (all {['%.3f' % v for v in ms]}) = {flops / best / 1e9:.2f} TF "
      f"({flops / best / 1e9 / 37.22496:.3f} of nominal); err vs zgemv {np.linalg.norm(Y[0] - y0) / np.linalg.norm(y0):.2e}")
