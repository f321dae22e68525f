"""bemb200_gmres_callback on one GPU: gmres_preconditioned with caller-supplied `Preconditioner` objects (host functions behind a
C function pointer) against the built-in device preconditioners and numpy.  Prints ONE JSON line; run by
tests/test_gpu_user_precond.py in its own process.  No oracle import: the comparison is between two paths of the library."""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[2]))
from math_audio_b200 import bem  # noqa: E402
from math_audio_b200.incident import IncidentField  # noqa: E402
from math_audio_b200.mesh import generate_icosphere_mesh  # noqa: E402
from math_audio_b200.types import PhysicsParams  # noqa: E402


class Counting:
    def __init__(self, fn):
        self.fn, self.calls = fn, 0

    def apply(self, r):
        self.calls += 1
        return self.fn(r)


def main():
    a = 0.1
    mesh = generate_icosphere_mesh(a, 2)  # 320 Tri3
    ph = PhysicsParams.from_wave_number(4.0 / a)
    beta, _ = ph.burton_miller_beta_adaptive(a)
    ctx = bem.Context(0)
    system = bem.build_tbem_system_with_beta(mesh, ph, beta, ctx=ctx)
    op = bem.DenseOperator(system)
    n = op.num_rows()
    A = system.matrix.rows()
    b = system.rhs + IncidentField.plane_wave_z().compute_rhs_with_beta(mesh.center, mesh.normal, ph, beta)
    cfg = bem.GmresConfig(max_iterations=200, restart=50, tolerance=1e-10)
    out = {"n": n}

    def true_res(x):
        return float(np.linalg.norm(b - A @ x) / np.linalg.norm(b))

    # (a) identity through the callback == built-in IdentityPreconditioner, bit for bit (the vector only makes a round trip)
    ref = bem.gmres_preconditioned(op, bem.IdentityPreconditioner(), b, cfg)
    ident = Counting(lambda r: r)
    sol = bem.gmres_preconditioned(op, ident, b, cfg)
    out["identity"] = {"iterations": [sol.iterations, ref.iterations], "restarts": [sol.restarts, ref.restarts],
                       "converged": bool(sol.converged and ref.converged), "max_abs_dx": float(np.max(np.abs(sol.x - ref.x))),
                       "calls": [ident.calls, sol.preconditioner_calls], "true_residual": true_res(sol.x)}
    # (b) Jacobi through the callback against the device's DiagonalPreconditioner
    dp = bem.DiagonalPreconditioner.from_operator(op)
    ref = bem.gmres_preconditioned(op, dp, b, cfg)
    jac = Counting(lambda r: r * dp.inv_diag)
    sol = bem.gmres_preconditioned(op, jac, b, cfg)
    out["jacobi"] = {"iterations": [sol.iterations, ref.iterations], "converged": bool(sol.converged and ref.converged),
                     "x_rel_diff": float(np.linalg.norm(sol.x - ref.x) / np.linalg.norm(ref.x)), "calls": jac.calls,
                     "true_residual": true_res(sol.x)}
    # (c) a preconditioner the library does not have: dense block inverses with LAPACK pivoting, small restart (several cycles)
    nb = 8
    edges = np.linspace(0, n, nb + 1).astype(int)
    inv = [np.linalg.inv(A[s:e, s:e]) for s, e in zip(edges[:-1], edges[1:])]

    def block(r):
        z = np.empty_like(r)
        for (s, e), m in zip(zip(edges[:-1], edges[1:]), inv):
            z[s:e] = m @ r[s:e]
        return z

    blk = Counting(block)
    cfg5 = bem.GmresConfig(max_iterations=200, restart=5, tolerance=1e-10)
    sol = bem.gmres_preconditioned(op, blk, b, cfg5)
    plain = bem.gmres_preconditioned(op, bem.IdentityPreconditioner(), b, cfg5)
    out["block"] = {"iterations": sol.iterations, "restarts": sol.restarts, "converged": bool(sol.converged), "calls": blk.calls,
                    "reported_calls": sol.preconditioner_calls, "plain_iterations": plain.iterations, "true_residual": true_res(sol.x)}
    # (d) with an initial guess: starting from the solution converges without an Arnoldi step
    sol0 = bem.gmres_preconditioned_with_guess(op, Counting(block), b, sol.x, bem.GmresConfig(max_iterations=200, restart=5, tolerance=1e-8))
    out["guess"] = {"iterations": sol0.iterations, "converged": bool(sol0.converged)}

    # (e) a failing callback ends the solve with its own exception; the operator stays usable
    class Boom:
        def apply(self, r):
            raise KeyError("user preconditioner failed")

    try:
        bem.gmres_preconditioned(op, Boom(), b, cfg)
        out["exception"] = "not raised"
    except KeyError as e:
        out["exception"] = f"KeyError: {e.args[0]}"

    class Short:
        def apply(self, r):
            return r[:-1]

    try:
        bem.gmres_preconditioned(op, Short(), b, cfg)
        out["wrong_shape"] = "not raised"
    except ValueError:
        out["wrong_shape"] = "ValueError"
    again = bem.gmres(op, b, cfg)
    out["after_failure"] = {"converged": bool(again.converged), "true_residual": true_res(again.x)}
    # the C function itself: non-zero return code -> BEMB200_ECALLBACK
    import ctypes as C

    from math_audio_b200 import _capi

    bad = _capi.PRECOND_FN(lambda u, r, z, nn: 7)
    x = np.empty(n, dtype=np.complex128)
    info = _capi.CGmresInfo()
    rc = _capi.lib().bemb200_gmres_callback(op.matrix._h, bad, None, _capi.ptr(b), None, 10, 10, 1e-10, _capi.ptr(x), C.byref(info), None)
    out["rc_on_nonzero_return"] = int(rc)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
