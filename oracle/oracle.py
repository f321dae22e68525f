"""ctypes front-end of the CPU parity oracle (oracle/bem_oracle.cpp).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(math_audio_b200) never imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_SO = _HERE / "_build" / "libbem_oracle.so"


def build(force: bool = False) -> Path:
    src = _HERE / "bem_oracle.cpp"
    if force or not _SO.exists() or _SO.stat().st_mtime < src.stat().st_mtime:
        subprocess.run(["make", "-C", str(_HERE)], check=True, stdout=subprocess.DEVNULL)
    return _SO


class _Mesh(C.Structure):
    _fields_ = [
        ("n_nodes", C.c_uint64), ("n_elem", C.c_uint64),
        ("nodes", C.c_void_p), ("conn", C.c_void_p), ("etype", C.c_void_p),
        ("center", C.c_void_p), ("normal", C.c_void_p), ("area", C.c_void_p),
        ("bc_type", C.c_void_p), ("bc_len", C.c_void_p), ("bc_val", C.c_void_p),
        ("dof", C.c_void_p), ("is_eval", C.c_void_p),
    ]


class GmresInfo(C.Structure):
    _fields_ = [("iterations", C.c_uint64), ("restarts", C.c_uint64), ("residual", C.c_double),
                ("converged", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(_SO))
        _lib.orc_assemble.restype = C.c_long
        _lib.orc_regular_integration.restype = C.c_long
        _lib.orc_singular_integration.restype = C.c_long
        _lib.orc_singular_integration_with_params.restype = C.c_long
        _lib.orc_dg_dn_sign.restype = C.c_double
        _lib.orc_count_dofs.restype = C.c_uint64
        _lib.orc_row_sum_correction.restype = C.c_double
        for f in ("orc_spherical_bessel_j", "orc_spherical_bessel_y", "orc_legendre_p", "orc_sphere_rcs"):
            getattr(_lib, f).restype = C.c_double
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _cmesh(mesh):
    """Keep references to the contiguous arrays alive on the returned struct."""
    arrs = dict(
        nodes=np.ascontiguousarray(mesh.nodes, dtype=np.float64),
        conn=np.ascontiguousarray(mesh.conn, dtype=np.uint32),
        etype=np.ascontiguousarray(mesh.etype, dtype=np.uint8),
        center=np.ascontiguousarray(mesh.center, dtype=np.float64),
        normal=np.ascontiguousarray(mesh.normal, dtype=np.float64),
        area=np.ascontiguousarray(mesh.area, dtype=np.float64),
        bc_type=np.ascontiguousarray(mesh.bc_type, dtype=np.int32),
        bc_len=np.ascontiguousarray(mesh.bc_len, dtype=np.uint8),
        bc_val=np.ascontiguousarray(mesh.bc_val, dtype=np.complex128),
        dof=np.ascontiguousarray(mesh.dof, dtype=np.uint32),
        is_eval=np.ascontiguousarray(mesh.is_eval, dtype=np.uint8),
    )
    m = _Mesh(mesh.n_nodes, mesh.n_elem, *[_p(arrs[k]) for k in
              ("nodes", "conn", "etype", "center", "normal", "area", "bc_type", "bc_len", "bc_val", "dof", "is_eval")])
    m._keep = arrs
    return m


def num_threads() -> int:
    return int(lib().orc_num_threads())


def gauss_legendre(order: int):
    n = C.c_int(0)
    x = np.zeros(32)
    w = np.zeros(32)
    lib().orc_gauss_legendre(C.c_int(order), C.byref(n), _p(x), _p(w))
    return x[: n.value].copy(), w[: n.value].copy()


def triangle_quadrature(order: int) -> np.ndarray:
    out = np.zeros((16, 3))
    n = lib().orc_triangle_quadrature(C.c_int(order), _p(out))
    return out[:n].copy()


def quad_quadrature(order: int) -> np.ndarray:
    out = np.zeros((400, 3))
    n = lib().orc_quad_quadrature(C.c_int(order), _p(out))
    return out[:n].copy()


def compute_parameters(coords, etype: int, s: float, t: float):
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    shape = np.zeros(4)
    jac = C.c_double(0)
    nrm = np.zeros(3)
    pos = np.zeros(3)
    lib().orc_compute_parameters(_p(coords), C.c_int(etype), C.c_double(s), C.c_double(t), _p(shape), C.byref(jac),
                                 _p(nrm), _p(pos))
    return shape[:etype].copy(), jac.value, nrm, pos


def local_to_global(coords, etype: int, s: float, t: float):
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    out = np.zeros(3)
    lib().orc_local_to_global(_p(coords), C.c_int(etype), C.c_double(s), C.c_double(t), _p(out))
    return out


def generate_subelements(src, coords, etype: int, area: float) -> np.ndarray:
    """rows: xi_c, eta_c, factor, gauss_order, v0xi, v0eta, v1xi, v1eta, v2xi, v2eta."""
    src = np.ascontiguousarray(src, dtype=np.float64)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    out = np.zeros((110, 10))
    n = lib().orc_generate_subelements(_p(src), _p(coords), C.c_int(etype), C.c_double(area), _p(out))
    return out[:n].copy()


_KEYS = ("g", "dg_dn", "dg_dnx", "d2g", "rhs")


def _unpack(o):
    return {k: complex(o[2 * i], o[2 * i + 1]) for i, k in enumerate(_KEYS)}


def regular_integration(src, nx, coords, etype, area, k, harmonic=1.0, tau=1.0, bc=None, bc_type=0, compute_rhs=False):
    src = np.ascontiguousarray(src, dtype=np.float64)
    nx = np.ascontiguousarray(nx, dtype=np.float64)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    out = np.zeros(10)
    bca = np.ascontiguousarray(bc, dtype=np.complex128) if bc is not None else None
    nq = lib().orc_regular_integration(_p(src), _p(nx), _p(coords), C.c_int(etype), C.c_double(area), C.c_double(k),
                                       C.c_double(harmonic), C.c_double(tau),
                                       _p(bca) if bca is not None else None, C.c_int(0 if bca is None else len(bca)),
                                       C.c_int(bc_type), C.c_int(1 if compute_rhs else 0), _p(out))
    r = _unpack(out)
    r["nqp"] = int(nq)
    return r


def singular_integration(src, nx, coords, etype, k, harmonic=1.0, tau=1.0, bc=None, bc_type=0, compute_rhs=False):
    src = np.ascontiguousarray(src, dtype=np.float64)
    nx = np.ascontiguousarray(nx, dtype=np.float64)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    out = np.zeros(10)
    bca = np.ascontiguousarray(bc, dtype=np.complex128) if bc is not None else None
    nq = lib().orc_singular_integration(_p(src), _p(nx), _p(coords), C.c_int(etype), C.c_double(k),
                                        C.c_double(harmonic), C.c_double(tau),
                                        _p(bca) if bca is not None else None, C.c_int(0 if bca is None else len(bca)),
                                        C.c_int(bc_type), C.c_int(1 if compute_rhs else 0), _p(out))
    r = _unpack(out)
    r["nqp"] = int(nq)
    return r


def singular_integration_with_params(src, nx, coords, etype, k, params, harmonic=1.0, tau=1.0):
    src = np.ascontiguousarray(src, dtype=np.float64)
    nx = np.ascontiguousarray(nx, dtype=np.float64)
    coords = np.ascontiguousarray(coords, dtype=np.float64)
    out = np.zeros(10)
    lib().orc_singular_integration_with_params(_p(src), _p(nx), _p(coords), C.c_int(etype), C.c_double(k),
                                               C.c_double(harmonic), C.c_double(tau), *[C.c_int(int(v)) for v in params],
                                               _p(out))
    return _unpack(out)


def dg_dn_sign(mesh, k: float) -> float:
    cm = _cmesh(mesh)
    return float(lib().orc_dg_dn_sign(C.byref(cm), C.c_double(k)))


def assemble(mesh, k, beta, row_begin=0, row_end=None, harmonic=1.0, tau=1.0, nthreads=0):
    """build_tbem_system_with_beta restricted to rows [row_begin,row_end) -> (A, rhs, n_qp)."""
    cm = _cmesh(mesh)
    ndof = mesh.num_dofs
    if row_end is None:
        row_end = ndof
    nr = row_end - row_begin
    A = np.empty((nr, ndof), dtype=np.complex128)
    rhs = np.empty(nr, dtype=np.complex128)
    beta = complex(beta)
    nq = lib().orc_assemble(C.byref(cm), C.c_double(k), C.c_double(harmonic), C.c_double(tau), C.c_double(beta.real),
                            C.c_double(beta.imag), C.c_uint64(row_begin), C.c_uint64(row_end), _p(A), _p(rhs),
                            C.c_int(nthreads))
    return A, rhs, int(nq)


def row_sum_correction(A: np.ndarray) -> float:
    assert A.flags.c_contiguous and A.dtype == np.complex128 and A.shape[0] == A.shape[1]
    return float(lib().orc_row_sum_correction(_p(A), C.c_uint64(A.shape[0])))


def zgemv(A: np.ndarray, x: np.ndarray, nthreads=0) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.complex128)
    x = np.ascontiguousarray(x, dtype=np.complex128)
    y = np.empty(A.shape[0], dtype=np.complex128)
    lib().orc_zgemv(_p(A), C.c_uint64(A.shape[0]), C.c_uint64(A.shape[1]), _p(x), _p(y), C.c_int(nthreads))
    return y


def zgemv_t(A: np.ndarray, x: np.ndarray) -> np.ndarray:
    A = np.ascontiguousarray(A, dtype=np.complex128)
    x = np.ascontiguousarray(x, dtype=np.complex128)
    y = np.empty(A.shape[1], dtype=np.complex128)
    lib().orc_zgemv_t(_p(A), C.c_uint64(A.shape[0]), C.c_uint64(A.shape[1]), _p(x), _p(y))
    return y


def gmres(A, b, x0=None, max_iterations=100, restart=30, tolerance=1e-6, nthreads=0):
    """gmres_with_guess on a dense row-major matrix -> (x, info dict)."""
    A = np.ascontiguousarray(A, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    n = b.shape[0]
    x = np.zeros(n, dtype=np.complex128)
    x0a = np.ascontiguousarray(x0, dtype=np.complex128) if x0 is not None else None
    info = GmresInfo()
    lib().orc_gmres(_p(A), C.c_uint64(n), _p(b), _p(x0a) if x0a is not None else None, C.c_uint32(max_iterations),
                    C.c_uint32(restart), C.c_double(tolerance), _p(x), C.byref(info), C.c_int(nthreads))
    return x, dict(iterations=int(info.iterations), restarts=int(info.restarts), residual=float(info.residual),
                   converged=bool(info.converged))


def inner_product(x, y) -> complex:
    """blas_helpers.rs:21-33: sum conj(x_i) y_i, sequential."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    y = np.ascontiguousarray(y, dtype=np.complex128)
    assert x.shape == y.shape
    out = np.zeros(2)
    lib().orc_inner_product(_p(x), _p(y), C.c_uint64(x.shape[0]), _p(out))
    return complex(out[0], out[1])


def vector_norm(x) -> float:
    """blas_helpers.rs:38-56: sqrt(sum |x_i|^2), sequential."""
    x = np.ascontiguousarray(x, dtype=np.complex128)
    lib().orc_vector_norm.restype = C.c_double
    return float(lib().orc_vector_norm(_p(x), C.c_uint64(x.shape[0])))


def bicgstab(A, b, max_iterations=1000, tolerance=1e-6, nthreads=0):
    """math-solvers/src/iterative/bicgstab.rs:46-187 on a dense row-major matrix -> (x, info dict)."""
    A = np.ascontiguousarray(A, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    n = b.shape[0]
    x = np.zeros(n, dtype=np.complex128)
    info = GmresInfo()
    lib().orc_bicgstab(_p(A), C.c_uint64(n), _p(b), C.c_uint32(max_iterations), C.c_double(tolerance), _p(x), C.byref(info),
                       C.c_int(nthreads))
    return x, dict(iterations=int(info.iterations), residual=float(info.residual), converged=bool(info.converged))


def cgs(A, b, max_iterations=1000, tolerance=1e-6, nthreads=0):
    """math-solvers/src/iterative/cgs.rs:46-155 on a dense row-major matrix -> (x, info dict)."""
    A = np.ascontiguousarray(A, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    n = b.shape[0]
    x = np.zeros(n, dtype=np.complex128)
    info = GmresInfo()
    lib().orc_cgs(_p(A), C.c_uint64(n), _p(b), C.c_uint32(max_iterations), C.c_double(tolerance), _p(x), C.byref(info),
                  C.c_int(nthreads))
    return x, dict(iterations=int(info.iterations), residual=float(info.residual), converged=bool(info.converged))


def lu_solve(A, b):
    """math-solvers/src/direct/lu.rs:139-161 -> x; raises np.linalg.LinAlgError for LuError::SingularMatrix."""
    A = np.ascontiguousarray(A, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    n = b.shape[0]
    if A.shape != (n, n):
        raise ValueError("dimension mismatch")
    x = np.zeros(n, dtype=np.complex128)
    lib().orc_lu_solve.restype = C.c_int
    if lib().orc_lu_solve(_p(A), C.c_uint64(n), _p(b), _p(x)) != 0:
        raise np.linalg.LinAlgError("Matrix is singular or nearly singular")
    return x


def inverse_diagonal(diag) -> np.ndarray:
    """DiagonalPreconditioner::from_diagonal (preconditioners/diagonal.rs:40-50): 1/d, or 1 when |d| <= 1e-30."""
    d = np.asarray(diag, dtype=np.complex128)
    out = np.ones_like(d)
    ok = np.sqrt(d.real ** 2 + d.imag ** 2) > 1e-30
    ns = d.real[ok] ** 2 + d.imag[ok] ** 2
    out[ok] = d.real[ok] / ns - 1j * (d.imag[ok] / ns)
    return out


def gmres_preconditioned(A, b, inv_diag=None, x0=None, max_iterations=100, restart=30, tolerance=1e-6, nthreads=0):
    """gmres_preconditioned_with_guess (gmres.rs:434-585) with the identity (inv_diag None) or the
    diagonal preconditioner -> (x, info dict)."""
    A = np.ascontiguousarray(A, dtype=np.complex128)
    b = np.ascontiguousarray(b, dtype=np.complex128)
    n = b.shape[0]
    x = np.zeros(n, dtype=np.complex128)
    x0a = np.ascontiguousarray(x0, dtype=np.complex128) if x0 is not None else None
    ida = np.ascontiguousarray(inv_diag, dtype=np.complex128) if inv_diag is not None else None
    info = GmresInfo()
    lib().orc_gmres_preconditioned(_p(A), C.c_uint64(n), _p(ida) if ida is not None else None, _p(b),
                                   _p(x0a) if x0a is not None else None, C.c_uint32(max_iterations), C.c_uint32(restart),
                                   C.c_double(tolerance), _p(x), C.byref(info), C.c_int(nthreads))
    return x, dict(iterations=int(info.iterations), restarts=int(info.restarts), residual=float(info.residual),
                   converged=bool(info.converged))


_APPLY_FN = C.CFUNCTYPE(None, C.c_void_p, C.c_void_p, C.c_void_p)


def gmres_preconditioned_cb(apply, precond_apply, n, b, x0=None, max_iterations=100, restart=30, tolerance=1e-6):
    """gmres_preconditioned_with_guess (gmres.rs:434-585) with operator AND preconditioner as Python callables
    (``apply(x) -> A x``, ``precond_apply(r) -> M^-1 r``), e.g. oracle/schwarz_oracle.py's restatement of schwarz.rs."""
    b = np.ascontiguousarray(b, dtype=np.complex128)
    x = np.zeros(n, dtype=np.complex128)
    x0a = np.ascontiguousarray(x0, dtype=np.complex128) if x0 is not None else None

    def _wrap(fn):
        def _cb(_user, xp, yp):
            xv = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)
            yv = np.ctypeslib.as_array(C.cast(yp, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)
            yv[:] = fn(xv.copy())
        return _APPLY_FN(_cb)

    cb_a, cb_p = _wrap(apply), _wrap(precond_apply)
    info = GmresInfo()
    lib().orc_gmres_preconditioned_cb(cb_a, None, cb_p, None, C.c_uint64(n), _p(b), _p(x0a) if x0a is not None else None,
                                      C.c_uint32(max_iterations), C.c_uint32(restart), C.c_double(tolerance), _p(x), C.byref(info))
    return x, dict(iterations=int(info.iterations), restarts=int(info.restarts), residual=float(info.residual),
                   converged=bool(info.converged))


def incident_rhs(kind, vec, amplitude, centers, normals, k, beta, tau=1.0):
    """compute_rhs_with_beta for one plane wave (kind=0) or point source (kind=1) -> (rhs, p_inc)."""
    centers = np.ascontiguousarray(centers, dtype=np.float64)
    normals = np.ascontiguousarray(normals, dtype=np.float64)
    vec = np.ascontiguousarray(vec, dtype=np.float64)
    n = centers.shape[0]
    rhs = np.zeros(n, dtype=np.complex128)
    pinc = np.zeros(n, dtype=np.complex128)
    amplitude = complex(amplitude)
    beta = complex(beta)
    lib().orc_incident_rhs(C.c_int(kind), _p(vec), C.c_double(amplitude.real), C.c_double(amplitude.imag), _p(centers),
                           _p(normals), C.c_uint64(n), C.c_double(k), C.c_double(tau), C.c_double(beta.real),
                           C.c_double(beta.imag), _p(rhs), _p(pinc), C.c_int(0))
    return rhs, pinc


def scattered_field(mesh, eval_points, surface_pressure, k, surface_velocity=None, harmonic=1.0, nthreads=0):
    cm = _cmesh(mesh)
    pts = np.ascontiguousarray(eval_points, dtype=np.float64)
    ps = np.ascontiguousarray(surface_pressure, dtype=np.complex128)
    vs = np.ascontiguousarray(surface_velocity, dtype=np.complex128) if surface_velocity is not None else None
    out = np.zeros(pts.shape[0], dtype=np.complex128)
    lib().orc_scattered_field(C.byref(cm), _p(pts), C.c_uint64(pts.shape[0]), _p(ps),
                              _p(vs) if vs is not None else None, C.c_double(k), C.c_double(harmonic), _p(out),
                              C.c_int(nthreads))
    return out


def compute_rcs(mesh, surface_pressure, directions, k):
    """postprocess/pressure.rs:438-478 for one or more unit directions."""
    cm = _cmesh(mesh)
    ps = np.ascontiguousarray(surface_pressure, dtype=np.complex128)
    dirs = np.ascontiguousarray(directions, dtype=np.float64).reshape(-1, 3)
    out = np.zeros(dirs.shape[0], dtype=np.float64)
    lib().orc_compute_rcs(C.byref(cm), _p(ps), _p(dirs), C.c_uint64(dirs.shape[0]), C.c_double(k), _p(out))
    return out


def mie_rigid_sphere(k, radius, num_terms, r, theta):
    r = np.ascontiguousarray(r, dtype=np.float64)
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    out = np.zeros(r.shape[0], dtype=np.complex128)
    lib().orc_mie_rigid_sphere(C.c_double(k), C.c_double(radius), C.c_int(num_terms), _p(r), _p(theta),
                               C.c_uint64(r.shape[0]), _p(out))
    return out


def spherical_bessel_j(n, x):
    return float(lib().orc_spherical_bessel_j(C.c_int(n), C.c_double(x)))


def spherical_bessel_y(n, x):
    return float(lib().orc_spherical_bessel_y(C.c_int(n), C.c_double(x)))


def legendre_p(n, x):
    return float(lib().orc_legendre_p(C.c_int(n), C.c_double(x)))


def sphere_rcs(k, radius, num_terms):
    return float(lib().orc_sphere_rcs(C.c_double(k), C.c_double(radius), C.c_int(num_terms)))


def l2_relative(analytical, bem) -> float:
    """ErrorMetrics::compute l2_relative: math-bem/src/testing/mod.rs:309-330."""
    a = np.asarray(analytical)
    b = np.asarray(bem)
    num = np.sqrt((np.abs(a - b) ** 2).sum())
    den = np.sqrt((np.abs(a) ** 2).sum())
    return float(num / den) if den > 1e-15 else float(num)


def gmres_op(apply, n, b, x0=None, max_iterations=100, restart=30, tolerance=1e-6):
    """gmres_with_guess (gmres.rs:105-277) over an arbitrary Python operator ``apply(x) -> y``
    (used to model the row-sharded operator of the multi-GPU path on CPU)."""
    b = np.ascontiguousarray(b, dtype=np.complex128)
    x = np.zeros(n, dtype=np.complex128)
    x0a = np.ascontiguousarray(x0, dtype=np.complex128) if x0 is not None else None

    def _cb(_user, xp, yp):
        xv = np.ctypeslib.as_array(C.cast(xp, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)
        yv = np.ctypeslib.as_array(C.cast(yp, C.POINTER(C.c_double)), shape=(2 * n,)).view(np.complex128)
        yv[:] = apply(xv.copy())

    cb = _APPLY_FN(_cb)
    info = GmresInfo()
    lib().orc_gmres_op(cb, None, C.c_uint64(n), _p(b), _p(x0a) if x0a is not None else None, C.c_uint32(max_iterations),
                       C.c_uint32(restart), C.c_double(tolerance), _p(x), C.byref(info))
    return x, dict(iterations=int(info.iterations), restarts=int(info.restarts), residual=float(info.residual),
                   converged=bool(info.converged))
