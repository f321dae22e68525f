"""Field evaluation and RCS on a mesh with a permuted DOF map and evaluation-only elements: the Python wrappers take
surface vectors as the reference does (entry j belongs to the j-th non-evaluation element, postprocess/pressure.rs:96-113,
452-458) and re-address them to the C ABI's DOF order; the result must be the oracle's for the same mesh and vectors.

The re-addressing itself is checked on the CPU (tests/test_solver_callers_cpu.py); this is its composition with the staged mesh
on the device (first green run on a B200: profiles/r02zz_pytest_dof_order.log)."""
import numpy as np
import pytest

from math_audio_b200.mesh import generate_icosphere_mesh
from math_audio_b200.types import PhysicsParams


@pytest.mark.gpu
def test_field_and_rcs_with_permuted_dofs_and_evaluation_elements(orc):
    from math_audio_b200 import bem

    a = 0.1
    mesh = generate_icosphere_mesh(a, 2)  # 320 Tri3
    rng = np.random.default_rng(21)
    mesh.is_eval[::11] = 1
    bnd = np.flatnonzero(mesh.is_eval == 0)
    nd = mesh.num_dofs
    mesh.dof[bnd] = rng.permutation(nd).astype(np.uint32)
    ph = PhysicsParams.from_wave_number(25.0)
    st = bem.StagedMesh(mesh)
    assert st.num_dofs == nd and st.enum_to_dof is not None
    p = rng.standard_normal(nd) + 1j * rng.standard_normal(nd)
    v = rng.standard_normal(nd) + 1j * rng.standard_normal(nd)
    pts = rng.standard_normal((41, 3))
    pts *= (3.0 * a / np.linalg.norm(pts, axis=1))[:, None]
    for vel in (None, v):
        got = bem.compute_scattered_field(pts, st, p, vel, ph)
        ref = orc.scattered_field(mesh, pts, p, ph.wave_number, surface_velocity=vel)
        assert np.linalg.norm(got - ref) / np.linalg.norm(ref) < 1e-12
    dirs = rng.standard_normal((8, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    ref = orc.compute_rcs(mesh, p, dirs, ph.wave_number)
    got = bem.compute_rcs(p, st, dirs, ph)
    assert np.max(np.abs(got - ref) / ref) < 1e-11
