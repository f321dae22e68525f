"""GPU parity of the callers of the hot path (bem_solver.rs): device BiCGSTAB, LU through cuSOLVER
and the BemSolver / BemSolution mirror, through the C ABI, against the oracle.

Tolerances: solutions relative 1e-8; BiCGSTAB iteration counts equal the oracle's on
these well-conditioned cases (tree vs sequential inner products change only the last bits)."""
import numpy as np
import pytest

from math_audio_b200.mesh import generate_icosphere_mesh
from math_audio_b200.types import PhysicsParams

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bem():
    from math_audio_b200 import bem as b

    b.default_context()
    return b


def test_bicgstab_kats_and_dense(bem, orc):
    A = np.array([[4, 1], [1, 3]], dtype=np.complex128)
    b = np.array([1, 2], dtype=np.complex128)
    sol = bem.bicgstab(bem.DenseOperator(A), b, bem.BiCgstabConfig(100, 1e-10, 0))       # bicgstab.rs:196-219
    assert sol.converged and np.linalg.norm(A @ sol.x - b) < 1e-8
    z = bem.bicgstab(bem.DenseOperator(A), np.zeros(2, dtype=complex), bem.BiCgstabConfig())
    assert z.converged and z.iterations == 0 and z.residual == 0.0 and not z.x.any()
    rng = np.random.default_rng(2)
    for n in (40, 777, 5000):
        A = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n) + 3 * np.eye(n)
        b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        op = bem.DenseOperator(A)
        sol = bem.bicgstab(op, b, bem.BiCgstabConfig(500, 1e-11, 0))
        xo, io = orc.bicgstab(A, b, max_iterations=500, tolerance=1e-11)
        assert sol.converged and io["converged"]
        assert sol.iterations == io["iterations"]
        assert np.linalg.norm(sol.x - xo) / np.linalg.norm(xo) < 1e-8
        assert abs(sol.residual - np.linalg.norm(A @ sol.x - b) / np.linalg.norm(b)) < 1e-9
        short = bem.bicgstab(op, b, bem.BiCgstabConfig(3, 1e-14, 0))
        assert not short.converged and short.iterations == 3


def test_cgs_kats_and_dense(bem, orc):
    """cgs.rs:46-155 through bemb200_cgs, and the fmm_interface.rs wrappers that reach it
    (solve_cgs :360, solve_with_ilu :389 -- which runs unpreconditioned CGS -- solve_tbem_with_ilu :441)."""
    A = np.array([[4, 1], [1, 3]], dtype=np.complex128)
    b = np.array([1, 2], dtype=np.complex128)
    sol = bem.cgs(bem.DenseOperator(A), b, bem.CgsConfig(100, 1e-10, 0))                 # cgs.rs:158-181
    assert sol.converged and np.linalg.norm(A @ sol.x - b) < 1e-8
    z = bem.cgs(bem.DenseOperator(A), np.zeros(2, dtype=complex), bem.CgsConfig())
    assert z.converged and z.iterations == 0 and z.residual == 0.0 and not z.x.any()
    rng = np.random.default_rng(5)
    for n in (40, 777, 5000):
        A = (rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))) / np.sqrt(n) + 3 * np.eye(n)
        b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
        op = bem.DenseOperator(A)
        sol = bem.solve_cgs(op, b, bem.CgsConfig(500, 1e-11, 0))
        xo, io = orc.cgs(A, b, max_iterations=500, tolerance=1e-11)
        assert sol.converged and io["converged"]
        assert sol.iterations == io["iterations"]
        assert np.linalg.norm(sol.x - xo) / np.linalg.norm(xo) < 1e-8
        assert abs(sol.residual - np.linalg.norm(A @ sol.x - b) / np.linalg.norm(b)) < 1e-8
        short = bem.cgs(op, b, bem.CgsConfig(3, 1e-14, 0))
        assert not short.converged and short.iterations == 3
        ilu = bem.solve_tbem_with_ilu(op, b, bem.CgsConfig(500, 1e-11, 0))
        assert ilu.iterations == sol.iterations and (ilu.x == sol.x).all()
    with pytest.raises(ValueError):
        bem.cgs(bem.DenseOperator(A), b[:-1], bem.CgsConfig())


def test_cgs_fmm_validation_cases(bem, orc):
    """math-bem/tests/test_fmm_validation.rs:246-300 (test_iterative_solver_with_operator), :480-530
    (test_solve_tbem_convenience), :714-785 (test_gmres_vs_cgs_convergence), :787-870 (robustness) on the device."""
    import math

    def tridiag(n, d, lo, up):
        A = np.zeros((n, n), dtype=np.complex128)
        for i in range(n):
            A[i, i] = d
            if i > 0:
                A[i, i - 1] = lo
            if i < n - 1:
                A[i, i + 1] = up
        return A

    for A, w, budget, fn in ((tridiag(10, 10.0, complex(-1.0, 0.1), complex(-1.0, -0.1)), 0.3, 200, bem.solve_cgs),
                             (tridiag(20, 10.0, complex(-1.0, 0.1), complex(-1.0, -0.1)), 0.3, 200, bem.solve_tbem_with_ilu),
                             (tridiag(20, 10.0, -1.0, -1.0), 0.25, 100, bem.solve_cgs)):
        n = A.shape[0]
        b = np.array([math.sin(i * w) for i in range(n)], dtype=np.complex128)
        op = bem.DenseOperator(A)
        sol = fn(op, b, bem.CgsConfig(budget, 1e-10, 0))
        xo, io = orc.cgs(A, b, max_iterations=budget, tolerance=1e-10)
        assert sol.converged and sol.iterations == io["iterations"]
        assert np.linalg.norm(b - A @ sol.x) / np.linalg.norm(b) < 1e-6
        assert np.linalg.norm(sol.x - xo) / np.linalg.norm(xo) < 1e-10
        g = bem.solve_gmres(op, b, bem.GmresConfig(100, 20, 1e-10))
        assert g.converged and np.linalg.norm(g.x - sol.x) / np.linalg.norm(g.x) < 1e-6
    A = tridiag(25, complex(6.0, 0.3), complex(-2.0, 0.1), complex(-1.5, -0.1))
    b = np.array([math.sin(i * 0.25) for i in range(25)], dtype=np.complex128)
    op = bem.DenseOperator(A)
    g = bem.solve_gmres(op, b, bem.GmresConfig(100, 25, 1e-10))
    c = bem.solve_cgs(op, b, bem.CgsConfig(100, 1e-10, 0))
    xo, io = orc.cgs(A, b, max_iterations=100, tolerance=1e-10)
    assert g.converged and np.linalg.norm(A @ g.x - b) / np.linalg.norm(b) < 1e-8
    assert c.converged == io["converged"] and abs(c.iterations - io["iterations"]) <= 1


def test_cgs_on_tbem_system(bem, orc):
    """CGS on an assembled TBEM matrix (what solve_tbem_with_ilu is called on), vs the oracle's CGS on the oracle's matrix."""
    from math_audio_b200.incident import IncidentField
    from math_audio_b200.mesh import generate_icosphere_mesh
    from math_audio_b200.types import PhysicsParams

    a = 0.1
    mesh = generate_icosphere_mesh(a, 3)
    ph = PhysicsParams.from_wave_number(1.5 / a)
    beta = ph.burton_miller_beta_adaptive(a)[0]
    system = bem.build_tbem_system_with_beta(mesh, ph, beta)
    b = system.rhs + IncidentField.plane_wave_z().compute_rhs_with_beta(mesh.center, mesh.normal, ph, beta)
    sol = bem.solve_with_ilu(system, b, bem.CgsConfig(1000, 1e-10, 0))
    Ao, _, _ = orc.assemble(mesh, ph.wave_number, beta)
    xo, io = orc.cgs(Ao, b, max_iterations=1000, tolerance=1e-10)
    assert sol.converged and io["converged"] and abs(sol.iterations - io["iterations"]) <= 1
    assert np.linalg.norm(sol.x - xo) / np.linalg.norm(xo) < 1e-8
    xg = bem.gmres(bem.DenseOperator(system), b, bem.GmresConfig(max_iterations=100, restart=50, tolerance=1e-11)).x
    assert np.linalg.norm(sol.x - xg) / np.linalg.norm(xg) < 1e-8


def test_lu_solve_kats_and_dense(bem, orc):
    A = np.array([[4 + 1j, 1], [1, 3 - 1j]])
    b = np.array([1 + 1j, 2 - 1j])
    assert np.max(np.abs(A @ bem.lu_solve(A, b) - b)) < 1e-10                            # lu.rs:178-196
    assert np.allclose(bem.lu_solve(np.eye(5), np.arange(1.0, 6.0)), np.arange(1.0, 6.0), atol=1e-10)
    with pytest.raises(bem.LuError):
        bem.lu_solve(np.array([[1.0, 2.0], [2.0, 4.0]]), np.array([1.0, 2.0]))           # lu.rs:211-219
    with pytest.raises(bem.LuError):
        bem.lu_solve(np.eye(3), np.ones(4))
    rng = np.random.default_rng(4)
    n = 1500
    A = rng.standard_normal((n, n)) + 1j * rng.standard_normal((n, n))
    b = rng.standard_normal(n) + 1j * rng.standard_normal(n)
    op = bem.DenseOperator(A)
    x = bem.lu_solve(op, b)
    assert np.linalg.norm(x - np.linalg.solve(A, b)) / np.linalg.norm(x) < 1e-9
    assert np.array_equal(op.matrix.rows(), A)           # the operator is left intact, as `lu_solve(&a, &b)`
    small = A[:200, :200].copy()
    assert np.linalg.norm(bem.lu_solve(small, b[:200]) - orc.lu_solve(small, b[:200])) / np.linalg.norm(b[:200]) < 1e-9


@pytest.mark.parametrize("method", ["Direct", "BiCgStab"])
def test_bem_solver_small_problem_and_field(bem, orc, method):
    """bem_solver.rs:654-686 (+ parity of the whole chain against the oracle)."""
    from math_audio_b200 import bem_solver as bs

    problem = bs.BemProblem.rigid_sphere_scattering_custom(0.1, 100.0, 343.0, 1.21, 4, 8)
    solver = bs.BemSolver.new().with_solver_method(getattr(bs.SolverMethod, method))
    sol = solver.solve(problem)
    assert sol.num_dofs() > 0 and sol.max_surface_pressure() > 0.0 and sol.mean_surface_pressure() > 0.0
    p = sol.evaluate_pressure([0.0, 0.0, 0.2])
    assert abs(p) > 0.0
    mesh = solver.prepare_elements(problem)
    ph = problem.physics
    beta = ph.burton_miller_beta_scaled(4.0)
    A, rhs0, _ = orc.assemble(mesh, ph.wave_number, beta)
    rhs, pinc = orc.incident_rhs(0, [0, 0, 1.0], 1.0, mesh.center, mesh.normal, ph.wave_number, beta)
    xo = orc.lu_solve(A, rhs0 + rhs) if method == "Direct" else orc.bicgstab(A, rhs0 + rhs, 1000, 1e-8)[0]
    assert np.linalg.norm(sol.surface_pressure - xo) / np.linalg.norm(xo) < (1e-8 if method == "Direct" else 1e-6)
    pts = np.array([[0.0, 0.0, 0.2], [0.3, -0.1, 0.05]])
    fps = sol.evaluate_pressure_field(pts)
    ps = orc.scattered_field(mesh, pts, sol.surface_pressure, ph.wave_number)
    assert max(abs(fp.p_scattered - q) for fp, q in zip(fps, ps)) < 1e-12
    assert abs(fps[0].p_total - p) == 0.0 and np.isfinite(fps[0].spl_db())


def test_bem_solver_rigid_sphere_mie(bem, orc):
    """Direct and BiCGSTAB agree with each other and with the Mie series as the QA suite expects (qa_suite.rs:175-179)."""
    from math_audio_b200 import bem_solver as bs

    problem = bs.BemProblem.rigid_sphere_scattering(0.1, 343.0 * 10.0 / (2 * np.pi), 343.0, 1.21)   # ka = 1 -> icosphere(3)
    assert problem.mesh.n_elem == 1280
    stats = {}
    xd = bs.BemSolver.new().solve(problem, stats=stats).surface_pressure
    assert stats["factor_ms"] > 0.0
    xb = bs.BemSolver.new().with_solver_method(bs.SolverMethod.BiCgStab).with_tolerance(1e-10).solve(problem).surface_pressure
    assert np.linalg.norm(xd - xb) / np.linalg.norm(xd) < 1e-8
    m = problem.mesh
    r = np.linalg.norm(m.center, axis=1)
    mie = orc.mie_rigid_sphere(problem.physics.wave_number, 0.1, 50, r, np.arccos(m.center[:, 2] / r))
    assert abs(orc.l2_relative(mie, xd) - 0.2724) < 2e-4


# ---- math-bem/tests/test_bem_sphere_integration.rs restated through the BemSolver mirror --------------------
def _field_vs_mie(bem_solver_mod, orc, frequency, n_theta, n_phi, n_pts, terms):
    radius, c0, rho = 0.1, 343.0, 1.21
    k = 2.0 * np.pi * frequency / c0
    problem = bem_solver_mod.BemProblem.rigid_sphere_scattering_custom(radius, frequency, c0, rho, n_theta, n_phi)
    solution = bem_solver_mod.BemSolver.new().solve(problem)
    er = 2.0 * radius
    thetas = np.pi * np.arange(n_pts) / (n_pts - 1)
    pts = np.column_stack([er * np.sin(thetas), np.zeros(n_pts), er * np.cos(thetas)])
    field = solution.evaluate_pressure_field(pts)
    mie = orc.mie_rigid_sphere(k, radius, terms, np.full(n_pts, er), thetas)
    rel = [abs(abs(fp.p_total) - abs(m)) / abs(m) if abs(m) > 1e-10 else 0.0 for fp, m in zip(field, mie)]
    return solution, max(rel)


def test_bem_vs_analytical_rayleigh(bem, orc):       # test_bem_sphere_integration.rs:23-118
    from math_audio_b200 import bem_solver as bs

    _, err = _field_vs_mie(bs, orc, 100.0, 6, 12, 9, 20)
    assert err < 0.5


def test_bem_vs_analytical_mie(bem, orc):            # :121-205
    from math_audio_b200 import bem_solver as bs

    _, err = _field_vs_mie(bs, orc, 546.0, 8, 16, 13, 30)
    assert err < 0.75


def test_bem_surface_pressure_distribution(bem, orc):  # :208-255
    from math_audio_b200 import bem_solver as bs

    sol = bs.BemSolver.new().solve(bs.BemProblem.rigid_sphere_scattering_custom(0.1, 200.0, 343.0, 1.21, 8, 16))
    assert sol.max_surface_pressure() > 0.0
    assert 1.0 < sol.max_surface_pressure() / sol.mean_surface_pressure() < 5.0


def test_bem_mesh_convergence_and_sanity(bem, orc):  # :262-346
    from math_audio_b200 import bem_solver as bs

    k = 2.0 * np.pi * 300.0 / 343.0
    er, th = 0.2, np.pi / 4
    mie = abs(orc.mie_rigid_sphere(k, 0.1, 30, np.array([er]), np.array([th]))[0])
    errs = []
    for nt, nphi in [(4, 8), (6, 12), (8, 16)]:
        sol = bs.BemSolver.new().solve(bs.BemProblem.rigid_sphere_scattering_custom(0.1, 300.0, 343.0, 1.21, nt, nphi))
        p = abs(sol.evaluate_pressure_field([[er * np.sin(th), 0.0, er * np.cos(th)]])[0].p_total)
        errs.append(abs(p - mie) / mie)
    assert min(errs) < 0.5
    sol = bs.BemSolver.new().solve(bs.BemProblem.rigid_sphere_scattering(0.1, 100.0, 343.0, 1.21))
    assert sol.num_dofs() > 0 and np.isfinite(sol.max_surface_pressure()) and sol.max_surface_pressure() > 0.0


# ---- math-bem/tests/test_accuracy_parity.rs restated (thresholds are the reference's) -----------------------
def _rel(c, r):                                      # test_accuracy_parity.rs:29-35
    return abs(c - r) / abs(r) if abs(r) > 1e-15 else abs(c)


def _solve_ka(bs, ka, n_theta, n_phi):
    radius, c0 = 0.1, 343.0
    k = ka / radius
    f = k * c0 / (2.0 * np.pi)
    return bs.BemSolver.new().with_solver_method(bs.SolverMethod.Direct).solve(
        bs.BemProblem.rigid_sphere_scattering_custom(radius, f, c0, 1.21, n_theta, n_phi)), k


def _ring(er, n):
    th = np.pi * np.arange(n + 1) / n
    return th, np.column_stack([er * np.sin(th), np.zeros(n + 1), er * np.cos(th)])


def test_accuracy_rayleigh_and_higher_frequency(bem, orc):   # :60-148, :266-326
    from math_audio_b200 import bem_solver as bs

    for ka, (nt, nphi), npts, terms, limit in [(0.1, (8, 16), 8, 30, 0.20), (0.2, (8, 16), 8, 30, 0.20), (0.3, (8, 16), 8, 30, 0.20),
                                               (2.0, (12, 24), 16, 50, 0.35)]:
        sol, k = _solve_ka(bs, ka, nt, nphi)
        th, pts = _ring(0.2, npts)
        mag = [abs(fp.p_total) for fp in sol.evaluate_pressure_field(pts)]
        mie = np.abs(orc.mie_rigid_sphere(k, 0.1, terms, np.full(len(th), 0.2), th))
        assert max(_rel(a, b) for a, b in zip(mag, mie)) < limit, ka


def test_accuracy_mie_regime_surface(bem, orc):              # :151-263
    from math_audio_b200 import bem_solver as bs

    for ka in (1.0, 1.2):
        sol, k = _solve_ka(bs, ka, 10, 20)
        th = np.pi * np.arange(13) / 12
        mie = np.abs(orc.mie_rigid_sphere(k, 0.1, 50, np.full(13, 0.1 * 1.001), th))
        c = sol.mesh.center
        elem_theta = np.arccos(c[:, 2] / np.linalg.norm(c, axis=1))
        worst = 0.0
        for t, ref in zip(th, mie):
            if ref < 0.1:
                continue
            j = int(np.argmin(np.abs(elem_theta - t)))       # first minimum, as the reference's strict `<` scan
            worst = max(worst, _rel(abs(sol.surface_pressure[j]), ref))
        assert worst < 0.30, ka


def test_mesh_convergence_forward_back_and_phase(bem, orc):  # :328-416, :419-497, :500-580
    from math_audio_b200 import bem_solver as bs

    k = 10.0
    er, th = 0.2, np.pi / 4
    ref = abs(orc.mie_rigid_sphere(k, 0.1, 50, np.array([er]), np.array([th]))[0])
    errs = []
    for nt, nphi in [(6, 12), (8, 16), (10, 20), (12, 24)]:
        sol, _ = _solve_ka(bs, 1.0, nt, nphi)
        errs.append(_rel(abs(sol.evaluate_pressure_field([[er * np.sin(th), 0.0, er * np.cos(th)]])[0].p_total), ref))
    assert errs[-1] < 0.25
    sol, _ = _solve_ka(bs, 1.0, 10, 20)
    f = sol.evaluate_pressure_field([[0.0, 0.0, 0.3], [0.0, 0.0, -0.3]])
    pf, pb = abs(f[0].p_total), abs(f[1].p_total)
    af = abs(orc.mie_rigid_sphere(k, 0.1, 40, np.array([0.3]), np.array([0.0]))[0])
    ab = abs(orc.mie_rigid_sphere(k, 0.1, 40, np.array([0.3]), np.array([np.pi]))[0])
    assert pf > 0.0 and pb > 0.0 and _rel(pf / pb, af / ab) < 0.50
    thr, pts = _ring(0.2, 8)
    mie = orc.mie_rigid_sphere(k, 0.1, 40, np.full(9, 0.2), thr)
    worst = 0.0
    for fp, m in zip(sol.evaluate_pressure_field(pts), mie):
        d = abs(np.angle(fp.p_total) - np.angle(m))
        worst = max(worst, 2 * np.pi - d if d > np.pi else d)
    assert worst < np.pi / 4


# ---- the boundary extensions of round 2: single-process multi-GPU group, sweep behind the C ABI ------------------------
@pytest.mark.parametrize("devices", [[0, 0], [0, 1], [0, 1, 2, 3]], ids=["one_gpu_twice", "two_gpus", "four_gpus"])
def test_single_process_group_two_ranks_on_one_gpu(bem, orc, devices):
    """bemb200_multi_*: one process, one rank context per listed device.  Listing device 0 twice splits its SMs between the two
    persistent solver kernels, so the row-sharded code path (peer-buffer exchange of the Krylov vector, cross-rank reduction
    round, sharded Gram-Schmidt) runs on the single GPU of the driver's test box; [0, 1] and [0, 1, 2, 3] are the real thing
    (cudaDeviceEnablePeerAccess between the devices; skipped when the box has fewer GPUs).  Entries, iteration counts,
    solution vs oracle."""
    import torch

    from math_audio_b200.incident import IncidentField

    if max(devices) >= torch.cuda.device_count():
        pytest.skip(f"needs {max(devices) + 1} GPUs")
    a = 0.1
    grp = bem.MultiGpu(devices)
    for sub, ka, restart in ((2, 1.0, 50), (3, 3.0, 50), (3, 8.0, 10)):
        mesh = generate_icosphere_mesh(a, sub)
        if sub == 3:
            mesh.is_eval[-7:] = 1  # n = 1273: uneven split
        n = mesh.num_dofs
        ph = PhysicsParams.from_wave_number(ka / a)
        beta, _ = ph.burton_miller_beta_adaptive(a)
        system = grp.build_tbem_system_with_beta(mesh, ph, beta)
        Ao, rhso, _ = orc.assemble(mesh, ph.wave_number, beta)
        A = system.rows()
        assert np.max(np.abs(A - Ao) / np.abs(Ao)) < 1e-10
        b = system.rhs + IncidentField.plane_wave_z().compute_rhs_with_beta(mesh.center[:n], mesh.normal[:n], ph, beta)
        cfg = bem.GmresConfig(max_iterations=30, restart=restart, tolerance=1e-10)
        sol = system.gmres(b, cfg)
        xo, io = orc.gmres(Ao, b, max_iterations=30, restart=restart, tolerance=1e-10)
        assert (sol.iterations, sol.restarts, sol.converged) == (io["iterations"], io["restarts"], io["converged"])
        assert np.linalg.norm(sol.x - xo) / np.linalg.norm(xo) < 1e-8
        system.close()
    grp.close()


def test_c_sweep_equals_sequential_solves(bem, orc):
    """bemb200_sweep_*: the pipelined schedule behind the C ABI returns, frequency by frequency, what
    build_tbem_system_with_beta + gmres return when called one after the other (iteration counts equal, x to 1e-9)."""
    from math_audio_b200.incident import IncidentField
    from math_audio_b200.sweep import Sweep

    a = 0.1
    mesh = generate_icosphere_mesh(a, 4)
    inc = IncidentField.plane_wave_z()
    cfg = bem.GmresConfig(max_iterations=1000, restart=50, tolerance=1e-10)
    cases = []
    for ka in (0.3, 1.0, 2.0, 4.0, 7.0):
        ph = PhysicsParams.from_wave_number(ka / a)
        beta, _ = ph.burton_miller_beta_adaptive(a)
        cases.append((ph, beta, inc.compute_rhs_with_beta(mesh.center, mesh.normal, ph, beta)))
    sw = Sweep(mesh)
    got = sw.solve_all(cases, cfg)
    # the same sweep with gmres_preconditioned + block-Jacobi (80 blocks of 64 DOFs) rebuilt from every frequency's matrix
    sw.set_block_jacobi(80)
    got_pc = sw.solve_all(cases, cfg)
    sw.set_block_jacobi(0)
    again = sw.solve_all(cases[:1], cfg)
    sw.close()
    assert again[0][0].iterations == got[0][0].iterations
    for (ph, beta, rhs), (sol, stats, b), (solp, _sp, _bp) in zip(cases, got, got_pc):
        system = bem.build_tbem_system_with_beta(mesh, ph, beta)
        op = bem.DenseOperator(system)
        ref = bem.gmres(op, system.rhs + rhs, cfg)
        assert np.array_equal(b, system.rhs + rhs)
        assert (sol.iterations, sol.restarts, sol.converged) == (ref.iterations, ref.restarts, ref.converged)
        assert np.linalg.norm(sol.x - ref.x) / np.linalg.norm(ref.x) < 1e-9
        assert stats["near_pairs"] > 0
        pre = bem.AdditiveSchwarzPreconditioner.from_operator(op, 80)
        refp = bem.gmres_preconditioned(op, pre, system.rhs + rhs, cfg)
        pre.close()
        assert (solp.iterations, solp.restarts, solp.converged) == (refp.iterations, refp.restarts, True)
        assert np.linalg.norm(solp.x - refp.x) / np.linalg.norm(refp.x) < 1e-9
        assert np.linalg.norm(solp.x - ref.x) / np.linalg.norm(ref.x) < 1e-8 and solp.iterations < sol.iterations
