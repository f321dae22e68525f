// A caller's own Preconditioner (math-solvers/src/traits.rs:366-371) through the C++ mirror: bemb200::Preconditioner is handed to
// the device solver as a host function (bemb200_gmres_callback).  Jacobi written by hand must behave like the built-in
// DiagonalPreconditioner, apply() is called exactly where the reference calls it (M^-1 b, once per restart cycle, once per Arnoldi
// step), and an exception thrown inside apply() comes out of gmres_preconditioned.  The matrix is the tridiagonal operator of
// math-bem/tests/test_fmm_validation.rs:537-700.  Built and run by tests/test_gpu_user_precond.py.
#include <cstdio>
#include <cstdlib>

#include "bemb200.hpp"

using namespace bemb200;

#define CHECK(cond)                                                                      \
    do {                                                                                 \
        if (!(cond)) {                                                                   \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                                \
        }                                                                                \
    } while (0)

struct HostJacobi : Preconditioner {
    std::vector<Complex64> inv;
    mutable std::size_t calls = 0;
    std::vector<Complex64> apply(const std::vector<Complex64>& r) const override {
        ++calls;
        std::vector<Complex64> z(r.size());
        for (std::size_t i = 0; i < r.size(); ++i) z[i] = r[i] * inv[i];
        return z;
    }
};
struct Thrower : Preconditioner {
    std::vector<Complex64> apply(const std::vector<Complex64>&) const override { throw std::runtime_error("user preconditioner failed"); }
};

int main() {
    Context ctx(0);
    const std::size_t n = 50;
    std::vector<Complex64> t(n * n, Complex64(0.0, 0.0)), ones(n, Complex64(1.0, 0.0));
    for (std::size_t i = 0; i < n; ++i) {
        t[i * n + i] = Complex64(4.0 + 0.1 * static_cast<double>(i % 7), 0.5);
        if (i + 1 < n) t[i * n + i + 1] = t[(i + 1) * n + i] = Complex64(-1.0, 0.0);
    }
    DenseOperator op(ctx, t, n, n);
    const GmresConfig cfg{100, 8, 1e-10, 0};
    GmresSolution jac = gmres_preconditioned(op, DiagonalPreconditioner::from_diagonal(op.diagonal()), ones, cfg);
    HostJacobi hj;
    hj.inv = DiagonalPreconditioner::from_diagonal(op.diagonal()).inv_diag;
    GmresSolution uj = gmres_preconditioned(op, hj, ones, cfg);
    CHECK(jac.converged && uj.converged);
    CHECK(uj.iterations + 1 >= jac.iterations && uj.iterations <= jac.iterations + 1);
    double num = 0.0, den = 0.0;
    for (std::size_t i = 0; i < n; ++i) { num += std::norm(uj.x[i] - jac.x[i]); den += std::norm(jac.x[i]); }
    CHECK(std::sqrt(num / den) < 1e-8);
    CHECK(hj.calls == uj.iterations + uj.restarts + 2);
    Thrower thrower;
    bool rethrown = false;
    try { gmres_preconditioned(op, thrower, ones, cfg); } catch (const std::runtime_error&) { rethrown = true; }
    CHECK(rethrown);
    CHECK(solve_gmres(op, ones, cfg).converged);  // the operator stays usable
    std::printf("PASS iterations=%zu restarts=%zu calls=%zu\n", uj.iterations, uj.restarts, hj.calls);
    return 0;
}
