// C++ host-side tests through include/bemb200.hpp (the C++ mirror of the reference's Rust API),
// written to read like the reference's own tests:
//   math-bem/src/core/assembly/tbem.rs:536-615           test_build_tbem_system (2-element mesh)
//   math-solvers/src/iterative/gmres.rs:631-705          test_gmres_simple / test_gmres_identity
//   math-bem/tests/test_fmm_validation.rs:537-700        test_gmres_with_operator / _restart_behavior
// plus entry / solution parity against the CPU oracle (linked: oracle/_build/libbem_oracle.so) on a
// UV sphere generated like math-bem/src/core/mesh/generators.rs:29-98.
// Built and run by tests/test_cpp_host_api.py (pytest -m gpu).
#include <cassert>
#include <cmath>
#include <cstdio>
#include <cstdlib>

#include "bemb200.hpp"

using namespace bemb200;

extern "C" {
struct orc_mesh {
    uint64_t n_nodes, n_elem;
    const double* nodes; const uint32_t* conn; const uint8_t* etype; const double* center; const double* normal; const double* area;
    const int32_t* bc_type; const uint8_t* bc_len; const double* bc_val; const uint32_t* dof; const uint8_t* is_eval;
};
struct orc_gmres_info { uint64_t iterations, restarts; double residual; int32_t converged; };
long orc_assemble(const orc_mesh* m, double k, double harmonic, double tau, double beta_re, double beta_im, uint64_t row_begin,
                  uint64_t row_end, double* A_out, double* rhs_out, int nthreads);
void orc_gmres(const double* A, uint64_t n, const double* b, const double* x0, uint32_t max_iterations, uint32_t restart,
               double tolerance, double* x_out, orc_gmres_info* info, int nthreads);
}

#define CHECK(cond)                                                              \
    do {                                                                         \
        if (!(cond)) {                                                           \
            std::fprintf(stderr, "CHECK failed %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            std::exit(1);                                                        \
        }                                                                        \
    } while (0)

static const double PI = 3.14159265358979323846264338327950288;

static void sphere_mesh(double radius, int n_theta, int n_phi, std::vector<double>& nodes, std::vector<Element>& elements) {
    nodes = {0.0, 0.0, radius};
    for (int i = 1; i < n_theta; ++i) {
        double theta = PI * i / n_theta;
        for (int j = 0; j < n_phi; ++j) {
            double phi = 2.0 * PI * j / n_phi;
            nodes.insert(nodes.end(), {radius * std::sin(theta) * std::cos(phi), radius * std::sin(theta) * std::sin(phi), radius * std::cos(theta)});
        }
    }
    nodes.insert(nodes.end(), {0.0, 0.0, -radius});
    const std::size_t south = nodes.size() / 3 - 1;
    std::vector<std::vector<std::size_t>> conn;
    for (int j = 0; j < n_phi; ++j) conn.push_back({0, (std::size_t)(1 + j), (std::size_t)(1 + (j + 1) % n_phi)});
    for (int i = 0; i < n_theta - 2; ++i) {
        std::size_t r0 = 1 + (std::size_t)i * n_phi, r1 = 1 + (std::size_t)(i + 1) * n_phi;
        for (int j = 0; j < n_phi; ++j) {
            std::size_t jn = (j + 1) % n_phi, n0 = r0 + j, n1 = r0 + jn, n2 = r1 + j, n3 = r1 + jn;
            conn.push_back({n0, n2, n1});
            conn.push_back({n1, n2, n3});
        }
    }
    std::size_t last = 1 + (std::size_t)(n_theta - 2) * n_phi;
    for (int j = 0; j < n_phi; ++j) conn.push_back({last + j, south, last + (std::size_t)((j + 1) % n_phi)});
    elements.clear();
    for (std::size_t e = 0; e < conn.size(); ++e) {  // create_mesh_from_data + compute_element_geometry
        Element el;
        el.connectivity = conn[e];
        el.element_type = ElementType::Tri3;
        const double* p0 = &nodes[3 * conn[e][0]]; const double* p1 = &nodes[3 * conn[e][1]]; const double* p2 = &nodes[3 * conn[e][2]];
        double v1[3], v2[3];
        for (int d = 0; d < 3; ++d) { el.center[d] = (p0[d] + p1[d] + p2[d]) / 3.0; v1[d] = p1[d] - p0[d]; v2[d] = p2[d] - p0[d]; }
        double cr[3] = {v1[1] * v2[2] - v1[2] * v2[1], v1[2] * v2[0] - v1[0] * v2[2], v1[0] * v2[1] - v1[1] * v2[0]};
        double len = std::sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
        el.area = len / 2.0;
        for (int d = 0; d < 3; ++d) el.normal[d] = cr[d] / len;
        if (el.normal[0] * el.center[0] + el.normal[1] * el.center[1] + el.normal[2] * el.center[2] < 0.0)
            for (int d = 0; d < 3; ++d) el.normal[d] = -el.normal[d];
        el.boundary_condition = BoundaryCondition::velocity({Complex64(0.0, 0.0)});
        el.dof_addresses = {e};
        elements.push_back(el);
    }
}

static std::vector<Complex64> tridiag(std::size_t n, Complex64 d, Complex64 lo, Complex64 up) {
    std::vector<Complex64> a(n * n, Complex64(0, 0));
    for (std::size_t i = 0; i < n; ++i) {
        a[i * n + i] = d;
        if (i > 0) a[i * n + i - 1] = lo;
        if (i + 1 < n) a[i * n + i + 1] = up;
    }
    return a;
}
static double rel_residual(const std::vector<Complex64>& a, std::size_t n, const std::vector<Complex64>& x, const std::vector<Complex64>& b) {
    double num = 0, den = 0;
    for (std::size_t i = 0; i < n; ++i) {
        Complex64 s(0, 0);
        for (std::size_t j = 0; j < n; ++j) s += a[i * n + j] * x[j];
        num += std::norm(s - b[i]); den += std::norm(b[i]);
    }
    return std::sqrt(num / den);
}

int main() {
    Context ctx(0);

    {  // tbem.rs:585-598 test_build_tbem_system
        std::vector<double> nodes = {0.0, 0.0, 0.0, 1.0, 0.0, 0.0, 0.5, 1.0, 0.0, 1.5, 1.0, 0.0};
        Element e0, e1;
        e0.connectivity = {0, 1, 2}; e0.normal[2] = 1.0; e0.center[0] = 0.5; e0.center[1] = 1.0 / 3.0; e0.area = 0.5;
        e0.boundary_condition = BoundaryCondition::velocity({Complex64(1.0, 0.0)}); e0.dof_addresses = {0};
        e1.connectivity = {1, 3, 2}; e1.normal[2] = 1.0; e1.center[0] = 1.0; e1.center[1] = 2.0 / 3.0; e1.area = 0.5;
        e1.boundary_condition = BoundaryCondition::velocity({Complex64(0.0, 0.0)}); e1.dof_addresses = {1};
        PhysicsParams physics(100.0, 343.0, 1.21, false);
        TbemSystem system = build_tbem_system(ctx, {e0, e1}, nodes, physics);
        CHECK(system.num_dofs == 2 && system.matrix->num_rows() == 2 && system.matrix->num_cols() == 2 && system.rhs.size() == 2);
        std::vector<Complex64> A = system.matrix->rows(0, 2);
        CHECK(std::abs(A[0]) > 1e-15 && std::abs(A[3]) > 1e-15);
    }
    {  // gmres.rs:631-680
        std::vector<Complex64> dense = {4.0, 1.0, 1.0, 3.0};
        DenseOperator a(ctx, dense, 2, 2);
        std::vector<Complex64> b = {1.0, 2.0};
        GmresSolution s = gmres(a, b, GmresConfig{100, 10, 1e-10, 0});
        CHECK(s.converged && rel_residual(dense, 2, s.x, b) < 1e-8);
        std::vector<Complex64> id(25, Complex64(0, 0)), bb(5);
        for (int i = 0; i < 5; ++i) { id[i * 5 + i] = 1.0; bb[i] = double(i + 1); }
        DenseOperator ident(ctx, id, 5, 5);
        s = gmres(ident, bb, GmresConfig{10, 10, 1e-12, 0});
        CHECK(s.converged && s.iterations <= 2);
        for (int i = 0; i < 5; ++i) CHECK(std::abs(s.x[i] - bb[i]) < 1e-10);
    }
    {  // test_fmm_validation.rs:537-700
        const std::size_t n = 20;
        auto m = tridiag(n, {10.0, 0.0}, {-1.0, 0.1}, {-1.0, -0.1});
        std::vector<Complex64> b(n);
        for (std::size_t i = 0; i < n; ++i) b[i] = std::sin(i * 0.3);
        DenseOperator op(ctx, m, n, n);
        GmresSolution s = solve_gmres(op, b, GmresConfig{50, 15, 1e-10, 0});
        CHECK(s.converged && rel_residual(m, n, s.x, b) < 1e-8);
        const std::size_t n2 = 50;
        auto t = tridiag(n2, {4.0, 0.0}, {-1.0, 0.0}, {-1.0, 0.0});
        std::vector<Complex64> ones(n2, Complex64(1.0, 0.0));
        DenseOperator op2(ctx, t, n2, n2);
        GmresSolution small = solve_gmres(op2, ones, GmresConfig{100, 5, 1e-10, 0});
        GmresSolution large = solve_gmres(op2, ones, GmresConfig{100, 50, 1e-10, 0});
        CHECK(small.converged && large.converged && large.restarts <= small.restarts);
        GmresSolution jac = gmres_preconditioned(op2, DiagonalPreconditioner::from_diagonal(op2.diagonal()), ones, GmresConfig{100, 50, 1e-10, 0});
        CHECK(jac.converged && rel_residual(t, n2, jac.x, ones) < 1e-8);
        bool threw = false;
        try { op2.apply(std::vector<Complex64>(7)); } catch (const std::invalid_argument&) { threw = true; }
        CHECK(threw);  // the reference panics on a shape mismatch
    }
    {  // rigid sphere, ka = 1, adaptive beta: all entries + GMRES against the oracle
        std::vector<double> nodes;
        std::vector<Element> elements;
        sphere_mesh(0.1, 12, 16, nodes, elements);
        const std::size_t n = elements.size();
        PhysicsParams physics(10.0 * 343.0 / (2.0 * PI), 343.0, 1.21, false);
        auto [beta, scale] = physics.burton_miller_beta_adaptive(0.1);
        CHECK(scale == 4.0);
        TbemSystem system = build_tbem_system_with_beta(ctx, elements, nodes, physics, beta);
        std::vector<Complex64> A = system.matrix->rows(0, n);
        // oracle on the same inputs
        std::vector<uint32_t> conn(4 * n, 0xFFFFFFFFu), dof(n);
        std::vector<uint8_t> etype(n, 3), bcl(n, 1), ev(n, 0);
        std::vector<double> cen(3 * n), nor(3 * n), area(n), bcv(8 * n, 0.0);
        std::vector<int32_t> bct(n, 0);
        for (std::size_t e = 0; e < n; ++e) {
            for (int v = 0; v < 3; ++v) conn[4 * e + v] = (uint32_t)elements[e].connectivity[v];
            for (int d = 0; d < 3; ++d) { cen[3 * e + d] = elements[e].center[d]; nor[3 * e + d] = elements[e].normal[d]; }
            area[e] = elements[e].area; dof[e] = (uint32_t)e;
        }
        orc_mesh om{nodes.size() / 3, n, nodes.data(), conn.data(), etype.data(), cen.data(), nor.data(), area.data(), bct.data(),
                    bcl.data(), bcv.data(), dof.data(), ev.data()};
        std::vector<Complex64> Ao(n * n), rhso(n);
        orc_assemble(&om, physics.wave_number, 1.0, 1.0, beta.real(), beta.imag(), 0, n, reinterpret_cast<double*>(Ao.data()),
                     reinterpret_cast<double*>(rhso.data()), 0);
        double worst = 0.0;
        for (std::size_t i = 0; i < n * n; ++i) worst = std::fmax(worst, std::abs(A[i] - Ao[i]) / std::abs(Ao[i]));
        CHECK(worst < 1e-10);
        // plane wave +z right-hand side (incident.rs:317-342) and solve
        std::vector<Complex64> b(n);
        const double k = physics.wave_number;
        for (std::size_t e = 0; e < n; ++e) {
            double kx = k * cen[3 * e + 2], kn = k * nor[3 * e + 2];
            Complex64 p(std::cos(kx), std::sin(kx));
            b[e] = -(p + beta * Complex64(0.0, kn) * p);
        }
        GmresSolution s = solve_gmres(*system.matrix, b, GmresConfig{1000, 50, 1e-10, 0});
        std::vector<Complex64> xo(n);
        orc_gmres_info info{};
        orc_gmres(reinterpret_cast<const double*>(Ao.data()), n, reinterpret_cast<const double*>(b.data()), nullptr, 1000, 50, 1e-10,
                  reinterpret_cast<double*>(xo.data()), &info, 0);
        double num = 0, den = 0;
        for (std::size_t e = 0; e < n; ++e) { num += std::norm(s.x[e] - xo[e]); den += std::norm(xo[e]); }
        CHECK(s.converged && s.iterations == info.iterations && std::sqrt(num / den) < 1e-8);
        std::printf("sphere N=%zu entry_err=%.2e gmres_it=%zu dx=%.2e\n", n, worst, s.iterations, std::sqrt(num / den));
    }
    std::printf("PASS\n");
    return 0;
}
