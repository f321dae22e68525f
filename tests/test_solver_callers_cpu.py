"""bicgstab / lu_solve / BemSolver, CPU side: the reference's own tests for these callers of the
hot path restated on the oracle, and the host logic of the BemSolver mirror.  No GPU."""
import math

import numpy as np
import pytest

from math_audio_b200 import bem_solver as bs


# ---- math-solvers/src/iterative/bicgstab.rs:196-219 -------------------------------------------
def test_bicgstab_simple(orc):
    A = np.array([[4, 1], [1, 3]], dtype=np.complex128)
    b = np.array([1, 2], dtype=np.complex128)
    x, info = orc.bicgstab(A, b, max_iterations=100, tolerance=1e-10)
    assert info["converged"]
    assert np.linalg.norm(A @ x - b) < 1e-8


def test_bicgstab_zero_rhs_and_budget(orc):
    A = np.array([[4, 1], [1, 3]], dtype=np.complex128)
    x, info = orc.bicgstab(A, np.zeros(2, dtype=np.complex128))
    assert info == dict(iterations=0, residual=0.0, converged=True) and not x.any()   # bicgstab.rs:66-74
    rng = np.random.default_rng(2)
    n = 40
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)) + 6 * np.eye(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x, info = orc.bicgstab(A, b, max_iterations=3, tolerance=1e-14)
    assert not info["converged"] and info["iterations"] == 3                           # bicgstab.rs:207-214
    x, info = orc.bicgstab(A, b, max_iterations=500, tolerance=1e-11)
    assert info["converged"] and np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-10
    assert abs(info["residual"] - np.linalg.norm(A @ x - b) / np.linalg.norm(b)) < 1e-12


# ---- math-solvers/src/iterative/cgs.rs:151-182 -------------------------------------------------
def test_cgs_simple(orc):
    A = np.array([[4, 1], [1, 3]], dtype=np.complex128)
    b = np.array([1, 2], dtype=np.complex128)
    x, info = orc.cgs(A, b, max_iterations=100, tolerance=1e-10)
    assert info["converged"]
    assert np.linalg.norm(A @ x - b) < 1e-8


def test_cgs_zero_rhs_budget_and_literal_restatement(orc):
    A = np.array([[4, 1], [1, 3]], dtype=np.complex128)
    x, info = orc.cgs(A, np.zeros(2, dtype=np.complex128))
    assert info == dict(iterations=0, residual=0.0, converged=True) and not x.any()   # cgs.rs:55-62
    rng = np.random.default_rng(3)
    n = 40
    A = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n) + 3 * np.eye(n)
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    x, info = orc.cgs(A, b, max_iterations=3, tolerance=1e-14)
    assert not info["converged"] and info["iterations"] == 3                           # cgs.rs:142-149
    x, info = orc.cgs(A, b, max_iterations=500, tolerance=1e-11)
    assert info["converged"] and np.linalg.norm(A @ x - b) / np.linalg.norm(b) < 1e-9

    # second, independent restatement of cgs.rs:64-140 in numpy (statement by statement)
    def cgs_np(A, b, max_iterations, tol):
        x = np.zeros_like(b)
        bn = np.linalg.norm(b)
        r = b.copy(); r0 = r.copy(); rho = np.vdot(r0, r); p = r.copy(); u = r.copy()
        for it in range(max_iterations):
            v = A @ p
            sigma = np.vdot(r0, v)
            if abs(sigma) < 1e-30:
                return x, it, False
            alpha = rho / sigma
            q = u - alpha * v
            upq = u + q
            w = A @ upq
            x = x + alpha * upq
            r = r - alpha * w
            rel = np.linalg.norm(r) / bn
            if rel < tol:
                return x, it + 1, True
            rho_new = np.vdot(r0, r)
            if abs(rho) < 1e-30:
                return x, it + 1, False
            beta = rho_new / rho
            rho = rho_new
            u = r + beta * q
            p = u + beta * (q + beta * p)
        return x, max_iterations, False

    x2, it2, conv2 = cgs_np(A, b, 500, 1e-11)
    assert conv2 and it2 == info["iterations"]
    assert np.linalg.norm(x - x2) / np.linalg.norm(x2) < 1e-9


# ---- math-solvers/src/direct/lu.rs:163-241 ----------------------------------------------------
def test_lu_solve_kats(orc):
    A = np.array([[4.0, 1.0], [1.0, 3.0]])
    b = np.array([1.0, 2.0])
    assert np.allclose(A @ orc.lu_solve(A, b), b, atol=1e-10)
    A = np.array([[4 + 1j, 1], [1, 3 - 1j]])
    b = np.array([1 + 1j, 2 - 1j])
    assert np.max(np.abs(A @ orc.lu_solve(A, b) - b)) < 1e-10
    assert np.allclose(orc.lu_solve(np.eye(5), np.arange(1.0, 6.0)), np.arange(1.0, 6.0), atol=1e-10)
    with pytest.raises(np.linalg.LinAlgError):
        orc.lu_solve(np.array([[1.0, 2.0], [2.0, 4.0]]), np.array([1.0, 2.0]))
    A = np.array([[4.0, 1.0, 0.0], [1.0, 3.0, 1.0], [0.0, 1.0, 2.0]])
    for b in (np.array([1.0, 2.0, 3.0]), np.array([4.0, 5.0, 6.0])):
        assert np.allclose(A @ orc.lu_solve(A, b), b, atol=1e-10)
    rng = np.random.default_rng(1)   # pivoting with a genuine 3-cycle permutation
    A = rng.standard_normal((30, 30)) + 1j * rng.standard_normal((30, 30))
    b = rng.standard_normal(30) + 1j * rng.standard_normal(30)
    assert np.linalg.norm(orc.lu_solve(A, b) - np.linalg.solve(A, b)) / np.linalg.norm(b) < 1e-11


# ---- math-bem/src/core/bem_solver.rs:633-651 --------------------------------------------------
def test_bem_problem_creation():
    p = bs.BemProblem.rigid_sphere_scattering(0.1, 1000.0, 343.0, 1.21)
    assert p.mesh.n_elem > 0 and p.mesh.n_nodes > 0 and p.ka() > 0.0
    assert p.mesh.n_elem == 1280                               # ka = 1.83 -> subdivisions 3 (bem_solver.rs:117-125)
    assert bs.BemProblem.rigid_sphere_scattering(0.1, 100.0, 343.0, 1.21).mesh.n_elem == 320
    assert bs.BemProblem.rigid_sphere_scattering(0.1, 5000.0, 343.0, 1.21).mesh.n_elem == 5120
    assert abs(p.ka() - 2 * math.pi * 1000.0 / 343.0 * 0.1) < 1e-9
    assert p.bc_type == bs.BoundaryConditionType.Rigid and p.use_burton_miller


def test_bem_solver_creation_and_prepare_elements():
    s = bs.BemSolver.new().with_solver_method(bs.SolverMethod.Direct).with_assembly_method(bs.AssemblyMethod.Tbem).with_verbose(False)
    assert s.solver_method == bs.SolverMethod.Direct and s.assembly_method == bs.AssemblyMethod.Tbem
    assert (s.max_iterations, s.tolerance, s.beta_scale) == (1000, 1e-8, 4.0)
    p = bs.BemProblem.rigid_sphere_scattering_custom(0.1, 100.0, 343.0, 1.21, 4, 8)
    m = s.prepare_elements(p)
    assert (m.bc_type == 0).all() and (m.bc_len == 1).all() and not m.bc_val.any() and (m.dof == np.arange(m.n_elem)).all()
    m = s.prepare_elements(p.with_boundary_condition(bs.BoundaryConditionType.Soft))
    assert (m.bc_type == 1).all()
    with pytest.raises(bs.BemError):
        bs.BemSolver.new().with_assembly_method(bs.AssemblyMethod.Mlfmm).solve(p)


# ---- math-bem/src/core/postprocess/pressure.rs:493-546 ----------------------------------------
def test_field_point_and_eval_point_generators():
    from math_audio_b200 import postprocess as pp

    fp = pp.FieldPoint(np.array([1.0, 0.0, 0.0]), 1.0 + 0j, 0.5 + 0.3j)
    assert abs(fp.p_total - (1.5 + 0.3j)) < 1e-10 and fp.magnitude() > 0.0 and np.isfinite(fp.spl_db())
    pts = pp.generate_sphere_eval_points(2.0, 10, 20)
    assert pts.shape == (200, 3) and np.allclose(np.linalg.norm(pts, axis=1), 2.0, atol=1e-10)
    line = pp.generate_line_eval_points([0.0, 0.0, 0.0], [1.0, 0.0, 0.0], 11)
    assert line.shape == (11, 3) and abs(line[0, 0]) < 1e-10 and abs(line[10, 0] - 1.0) < 1e-10 and abs(line[1, 0] - line[0, 0] - 0.1) < 1e-10
    assert pp.generate_line_eval_points([1.0, 2.0, 3.0], [4.0, 5.0, 6.0], 1).tolist() == [[1.0, 2.0, 3.0]]   # (n-1).max(1)
    plane = pp.generate_plane_eval_points([0.0, 0.0, 0.0], [0.0, 0.0, 1.0], 1.0, 5)
    assert plane.shape == (25, 3) and np.all(np.abs(plane[:, 2]) < 1e-10)
    tilted = pp.generate_plane_eval_points([1.0, 1.0, 1.0], [2.0, 0.0, 0.0], 0.5, 3)          # |n.x| >= 0.9 branch
    assert np.allclose(tilted[:, 0], 1.0) and np.allclose(tilted.mean(axis=0), [1.0, 1.0, 1.0])


# ---- user Preconditioner behind bemb200_precond_fn: the Python trampoline, driven by a stand-in for the C side -------------------
def test_user_preconditioner_trampoline_marshals_vectors_and_exceptions(monkeypatch, orc):
    """`gmres_preconditioned(op, any_object_with_apply, ...)` hands `apply` to the library as a C function pointer.  Here the
    library call is replaced by a host stand-in with the same signature that runs the oracle's preconditioned GMRES and calls the
    function pointer exactly as csrc/gmres.cu does (pointers to interleaved doubles, n, return code): the trampoline must
    marshal r and z correctly, count calls, and carry exceptions out of the C frames."""
    import ctypes as C

    from math_audio_b200 import _capi, bem

    rng = np.random.default_rng(5)
    n = 40
    A = np.eye(n) * 4.0 + 0.3 * (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n)))
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)

    class FakeLib:
        def bemb200_gmres_callback(self, mh, cb, user, b_ptr, x0_ptr, max_iterations, restart, tol, x_ptr, info_ref, calls_ref):
            bb = np.ctypeslib.as_array(C.cast(b_ptr, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)
            rbuf, zbuf = (C.c_double * (2 * n))(), (C.c_double * (2 * n))()
            state = {"calls": 0, "rc": 0}

            def precond(r):
                if state["rc"]:  # the real library stops at the first failure; the oracle's loop cannot be left from a callback
                    return np.zeros_like(r)
                np.frombuffer(rbuf, dtype=np.complex128)[:] = r
                state["calls"] += 1
                rc = cb(user, C.cast(rbuf, C.POINTER(C.c_double)), C.cast(zbuf, C.POINTER(C.c_double)), n)
                if rc != 0:
                    state["rc"] = rc
                    return np.zeros_like(r)
                return np.frombuffer(zbuf, dtype=np.complex128).copy()

            x, info = orc.gmres_preconditioned_cb(lambda v: A @ v, precond, n, bb.copy(), max_iterations=max_iterations,
                                                  restart=restart, tolerance=tol)
            if state["rc"]:
                C.cast(calls_ref, C.POINTER(C.c_uint64))[0] = state["calls"]
                return -8
            np.ctypeslib.as_array(C.cast(x_ptr, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)[:] = x
            inf = C.cast(info_ref, C.POINTER(_capi.CGmresInfo))[0]
            inf.iterations, inf.restarts, inf.residual, inf.converged = info["iterations"], info["restarts"], info["residual"], int(info["converged"])
            C.cast(calls_ref, C.POINTER(C.c_uint64))[0] = state["calls"]
            return 0

        def bemb200_last_error(self, ctx):
            return b"the preconditioner callback returned a non-zero code"

    class FakeMatrix:
        _h = None
        shape = (n, n)

        class ctx:
            _h = None

    op = bem.DenseOperator.__new__(bem.DenseOperator)
    op.matrix = FakeMatrix()
    monkeypatch.setattr(_capi, "lib", lambda: FakeLib())

    class Jacobi:
        def __init__(self):
            self.seen = 0

        def apply(self, r):
            assert r.dtype == np.complex128 and r.shape == (n,)
            self.seen += 1
            return r / np.diag(A)

    jac = Jacobi()
    cfg = bem.GmresConfig(max_iterations=50, restart=7, tolerance=1e-12)
    sol = bem.gmres_preconditioned(op, jac, b, cfg)
    xo, io = orc.gmres_preconditioned_cb(lambda v: A @ v, lambda r: r / np.diag(A), n, b, max_iterations=50, restart=7, tolerance=1e-12)
    assert sol.converged and sol.iterations == io["iterations"] and sol.restarts == io["restarts"]
    assert np.array_equal(sol.x, xo)
    assert jac.seen == sol.preconditioner_calls == sol.iterations + sol.restarts + 2
    assert np.linalg.norm(A @ sol.x - b) / np.linalg.norm(b) < 1e-10

    class Boom:
        def apply(self, r):
            raise KeyError("user preconditioner failed")

    with pytest.raises(KeyError):
        bem.gmres_preconditioned(op, Boom(), b, cfg)

    class Short:
        def apply(self, r):
            return r[:-1]

    with pytest.raises(ValueError):
        bem.gmres_preconditioned(op, Short(), b, cfg)
    with pytest.raises(TypeError):
        bem.gmres_preconditioned(op, object(), b, cfg)


# ---- surface vectors of a mesh with a permuted DOF map: enumeration order (reference) -> DOF order (C ABI) ------------------------
def test_surface_values_are_paired_with_elements_as_the_reference_pairs_them(orc):
    """postprocess/pressure.rs:96-113, 452-458 pair entry j of a surface vector with the j-th non-evaluation element; the device
    kernels walk the staged mesh in DOF order and pair an element with the entry at its DOF address.  The host wrappers
    re-address the vector (bem.surface_values_in_dof_order).  Checked with the oracle alone: evaluating the permuted mesh with
    the caller's vector == evaluating the DOF-sorted, sequentially numbered mesh (what the device walks) with the re-addressed
    vector -- field and RCS, with evaluation-only elements in between."""
    import dataclasses

    from math_audio_b200 import bem
    from math_audio_b200.mesh import generate_icosphere_mesh

    mesh = generate_icosphere_mesh(0.1, 1)  # 80 Tri3
    assert bem.enumeration_to_dof(mesh) is None  # generators number sequentially: nothing to do
    rng = np.random.default_rng(11)
    mesh.is_eval[::9] = 1
    nd = mesh.num_dofs
    bnd = np.flatnonzero(mesh.is_eval == 0)
    mesh.dof[bnd] = rng.permutation(nd).astype(np.uint32)
    e2d = bem.enumeration_to_dof(mesh)
    assert e2d is not None and np.array_equal(e2d, mesh.dof[bnd])
    p = rng.standard_normal(nd) + 1j * rng.standard_normal(nd)
    v = rng.standard_normal(nd) + 1j * rng.standard_normal(nd)
    p_dof, v_dof = bem.surface_values_in_dof_order(e2d, p), bem.surface_values_in_dof_order(e2d, v)
    for j, e in enumerate(bnd):
        assert p_dof[mesh.dof[e]] == p[j] and v_dof[mesh.dof[e]] == v[j]
    # the mesh the device walks: boundary elements sorted by DOF address, numbered 0..nd-1, no evaluation elements
    order = bnd[np.argsort(mesh.dof[bnd])]
    fields = {f.name: getattr(mesh, f.name) for f in dataclasses.fields(mesh)}
    for name in ("conn", "etype", "center", "normal", "area", "bc_type", "bc_len", "bc_val", "dof", "is_eval"):
        fields[name] = np.ascontiguousarray(fields[name][order])
    staged_view = type(mesh)(**fields)
    staged_view.dof[:] = np.arange(nd, dtype=np.uint32)
    k = 12.0
    pts = 0.3 * rng.standard_normal((9, 3)) + np.array([0.0, 0.0, 0.5])
    a = orc.scattered_field(mesh, pts, p, k, surface_velocity=v)
    b = orc.scattered_field(staged_view, pts, p_dof, k, surface_velocity=v_dof)
    assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(a))
    dirs = rng.standard_normal((5, 3))
    dirs /= np.linalg.norm(dirs, axis=1)[:, None]
    ra, rb = orc.compute_rcs(mesh, p, dirs, k), orc.compute_rcs(staged_view, p_dof, dirs, k)
    assert np.max(np.abs(ra - rb)) <= 1e-12 * np.max(np.abs(ra))
    # a map that is not a permutation is left to bemb200_mesh_stage to refuse
    mesh.dof[bnd[0]] = mesh.dof[bnd[1]]
    assert bem.enumeration_to_dof(mesh) is None
