"""Unit tests of the reference for the small pieces either side of the hot path, restated on the
oracle and on the host-side mirror (CPU only):

    math-bem/src/core/incident.rs:355-424            plane wave / point source / normal derivative / rhs
    math-bem/src/core/types.rs:741-767               PhysicsParams::new, BoundaryCondition indices
    math-bem/src/core/mesh/element.rs:252-318        Tri3 shape functions, local_to_global, normal, area
    math-solvers/src/blas_helpers.rs:146-256         inner_product (conjugates x), vector_norm, axpy family
    math-solvers/src/traits.rs:428-435               IdentityPreconditioner
    math-solvers/src/preconditioners/diagonal.rs:105-145   DiagonalPreconditioner
"""
import math

import numpy as np

from math_audio_b200 import bem
from math_audio_b200.incident import IncidentField
from math_audio_b200.mesh import mesh_from_data
from math_audio_b200.types import PhysicsParams


def make_physics(k):  # incident.rs:348-353
    c = 343.0
    return PhysicsParams.new(k * c / (2.0 * math.pi), c, 1.21, False)


# ---- incident.rs:355-424 ---------------------------------------------------------------------------
def test_plane_wave_on_axis(orc):
    inc = IncidentField.plane_wave([0.0, 0.0, 1.0], 1.0)
    ph = make_physics(1.0)
    pts = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 1.0], [0.0, 0.0, -1.0]])
    p = inc.evaluate_pressure(pts, ph)
    assert abs(p[0].real - 1.0) < 1e-10 and abs(p[0].imag) < 1e-10
    assert abs(p[1].real - math.cos(1.0)) < 1e-10 and abs(p[1].imag - math.sin(1.0)) < 1e-10
    # the oracle's restatement of the same function (it returns p_inc beside the rhs)
    _, po = orc.incident_rhs(0, [0.0, 0.0, 1.0], 1.0, pts, np.tile([0.0, 0.0, 1.0], (3, 1)), ph.wave_number, 0j)
    assert np.max(np.abs(po - p)) < 1e-15


def test_point_source_decay(orc):
    inc = IncidentField.point_source([0.0, 0.0, 0.0], 1.0)
    ph = make_physics(1.0)
    pts = np.array([[1.0, 0.0, 0.0], [2.0, 0.0, 0.0], [4.0, 0.0, 0.0]])
    p = inc.evaluate_pressure(pts, ph)
    assert abs(abs(p[0]) / abs(p[1]) - 2.0) < 0.1 and abs(abs(p[1]) / abs(p[2]) - 2.0) < 0.1
    _, po = orc.incident_rhs(1, [0.0, 0.0, 0.0], 1.0, pts, np.tile([1.0, 0.0, 0.0], (3, 1)), ph.wave_number, 0j)
    assert np.max(np.abs(po - p)) < 1e-15


def test_plane_wave_normal_derivative():
    inc = IncidentField.plane_wave([0.0, 0.0, 1.0], 1.0)
    ph = make_physics(1.0)
    d = inc.evaluate_normal_derivative(np.zeros((1, 3)), np.array([[0.0, 0.0, 1.0]]), ph)
    assert abs(d[0].real) < 1e-10 and abs(d[0].imag - 1.0) < 1e-10        # ik (d.n) p = +i at k = 1


def test_rhs_computation(orc):
    inc = IncidentField.plane_wave([0.0, 0.0, 1.0], 1.0)
    ph = make_physics(1.0)
    c = np.array([[0.0, 0.0, 1.0]])
    n = np.array([[0.0, 0.0, 1.0]])
    rhs = inc.compute_rhs(c, n, ph, False)
    assert abs(rhs[0]) > 0.0
    assert abs(rhs[0] + inc.evaluate_pressure(c, ph)[0]) < 1e-15           # -gamma p_inc without Burton-Miller
    bm = inc.compute_rhs(c, n, ph, True)
    ref, _ = orc.incident_rhs(0, [0.0, 0.0, 1.0], 1.0, c, n, ph.wave_number, ph.burton_miller_beta())
    assert abs(bm[0] - ref[0]) < 1e-15
    # direction is normalised, a zero direction falls back to -z (incident.rs:62-76)
    assert np.allclose(IncidentField.plane_wave([0.0, 3.0, 4.0]).plane_waves[0][0], [0.0, 0.6, 0.8])
    assert np.allclose(IncidentField.plane_wave([0.0, 0.0, 0.0]).plane_waves[0][0], [0.0, 0.0, -1.0])


# ---- types.rs:741-767 ------------------------------------------------------------------------------
def test_physics_params():
    p = PhysicsParams.new(1000.0, 343.0, 1.21, False)
    assert abs(p.wave_number - 2.0 * math.pi * 1000.0 / 343.0) < 1e-10
    assert abs(p.wave_length - 0.343) < 1e-10
    assert p.tau == 1.0 and PhysicsParams.new(1000.0, 343.0, 1.21, True).tau == -1.0
    assert p.harmonic_factor == 1.0 and p.gamma() == 1.0
    assert abs(p.pressure_factor - 1.21 * 2.0 * math.pi * 1000.0) < 1e-9


def test_boundary_condition_type_indices():
    # BoundaryCondition::type_index (types.rs:278-292) as the ABI carries it: 0 velocity, 1 pressure, 2 transfer
    nodes = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    m = mesh_from_data(nodes, [[0, 1, 2]])
    assert m.bc_type[0] == 0 and m.bc_len[0] == 1 and not m.bc_val.any()   # default: rigid, Velocity([0])
    m.set_velocity_bc([1.0 + 0j])
    assert m.bc_type[0] == 0 and m.bc_val.reshape(-1)[0] == 1.0
    assert m.etype[0] == 3                                                 # ElementType::Tri3.num_nodes()


# ---- element.rs:252-318 (the oracle's compute_parameters is regular.rs:193-260; same unit triangle) ----
def test_unit_triangle_shape_functions_normal_area(orc):
    tri = np.array([[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0]])
    for s, t in ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0), (1.0 / 3.0, 1.0 / 3.0), (0.5, 0.25)):
        shape, jac, normal, pos = orc.compute_parameters(tri, 3, s, t)
        assert abs(sum(shape[:3]) - 1.0) < 1e-12
        assert abs(jac - 1.0) < 1e-12 and np.allclose(normal, [0.0, 0.0, 1.0], atol=1e-12)   # 2 x area = 1
        assert np.allclose(pos, orc.local_to_global(tri, 3, s, t), atol=1e-15)
    # the three vertices are reached by the three unit shape vectors; the centroid by (1/3, 1/3, 1/3)
    hit = set()
    for s, t in ((0.0, 0.0), (1.0, 0.0), (0.0, 1.0)):
        shape, _, _, pos = orc.compute_parameters(tri, 3, s, t)
        i = int(np.argmax(shape[:3]))
        assert abs(shape[i] - 1.0) < 1e-12 and np.allclose(pos, tri[i], atol=1e-12)
        hit.add(i)
    assert hit == {0, 1, 2}
    shape, _, _, pos = orc.compute_parameters(tri, 3, 1.0 / 3.0, 1.0 / 3.0)
    assert np.allclose(shape[:3], 1.0 / 3.0, atol=1e-12) and np.allclose(pos, tri.mean(axis=0), atol=1e-12)
    m = mesh_from_data(tri, [[0, 1, 2]])
    assert abs(m.area[0] - 0.5) < 1e-10                                    # compute_element_area


# ---- blas_helpers.rs:146-256 -----------------------------------------------------------------------
def test_inner_product_and_norms(orc):
    assert orc.inner_product([1.0, 2.0, 3.0], [4.0, 5.0, 6.0]) == 32.0
    ip = orc.inner_product([1 + 2j, 3 + 4j], [5 + 6j, 7 + 8j])            # conj(x) . y
    assert abs(ip.real - 70.0) < 1e-10 and abs(ip.imag + 8.0) < 1e-10
    assert abs(orc.vector_norm([3.0, 4.0]) - 5.0) < 1e-10
    assert abs(orc.vector_norm([3.0 + 0j, 4j]) - 5.0) < 1e-10
    assert orc.vector_norm([0.0, 0.0, 0.0]) == 0.0
    assert abs(orc.vector_norm([3.0, 4.0]) ** 2 - 25.0) < 1e-10           # vector_norm_sqr
    rng = np.random.default_rng(7)
    x = rng.standard_normal(1000) + 1j * rng.standard_normal(1000)
    y = rng.standard_normal(1000) + 1j * rng.standard_normal(1000)
    assert abs(orc.inner_product(x, y) - np.vdot(x, y)) < 1e-11
    assert abs(orc.vector_norm(x) - np.linalg.norm(x)) < 1e-11


# ---- traits.rs:428-435, diagonal.rs:105-145 --------------------------------------------------------
def test_identity_and_diagonal_preconditioner(orc):
    r = np.array([1.0, 2.0, 3.0], dtype=np.complex128)
    z = bem.IdentityPreconditioner().apply(r)
    assert (z == r).all() and z is not r
    pre = bem.DiagonalPreconditioner.from_diagonal(np.array([2.0, 4.0, 1.0], dtype=np.complex128))
    out = pre.apply(np.array([2.0, 8.0, 3.0], dtype=np.complex128))
    assert np.allclose(out.real, [1.0, 2.0, 3.0], atol=1e-10) and not out.imag.any()
    A = np.array([[4.0, 1.0], [1.0, 2.0]], dtype=np.complex128)             # from_csr takes the diagonal
    out = bem.DiagonalPreconditioner.from_diagonal(np.diag(A)).apply(np.array([4.0, 4.0], dtype=np.complex128))
    assert np.allclose(out.real, [1.0, 2.0], atol=1e-10)
    assert np.allclose(pre.inv_diag, orc.inverse_diagonal(np.array([2.0, 4.0, 1.0])))


# ---- fmm_interface.rs:543-602 (mesh sizing helpers beside DenseOperator) ---------------------------------------------
def test_mesh_sizing_helpers():
    assert abs(bem.recommended_mesh_resolution(343.0, 343.0, 6) - 6.0) < 1e-12          # wavelength 1 m
    assert abs(bem.recommended_mesh_resolution(1000.0, 343.0, 8) - 8.0 / 0.343) < 1e-12
    assert bem.mesh_resolution_for_frequency_range(20.0, 500.0, 343.0, 6) == bem.recommended_mesh_resolution(500.0, 343.0, 6)
    assert bem.estimate_element_count((5.0, 4.0, 3.0), 2.0) == 376                      # 94 m^2 / 0.25 m^2
    assert bem.estimate_element_count((1.0, 1.0, 1.0), 1.5) == 14                       # ceil(6 * 2.25)
    cfg = bem.AdaptiveMeshConfig.for_frequency_range(20.0, 200.0)
    assert abs(cfg.base_resolution - 6.0 * 200.0 / 343.0) < 1e-12 and (cfg.source_refinement, cfg.source_refinement_radius) == (1.5, 0.5)
    fixed = bem.AdaptiveMeshConfig.from_resolution(4.0)
    assert (fixed.base_resolution, fixed.source_refinement, fixed.source_refinement_radius) == (4.0, 1.0, 0.0)
