// Shared host/device helpers of libbemb200 (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace bemb {

// device (__constant__) copy of the quadrature tables ...
#define BEMQ_TABLE_QUAL static __device__ __constant__ const
#include "quad_tables.h"
#undef BEMQ_TABLE_QUAL
// ... and a host copy for table set-up code
namespace hosttab {
#define BEMQ_TABLE_QUAL static const
#include "quad_tables.h"
#undef BEMQ_TABLE_QUAL
}  // namespace hosttab

constexpr double PI = 3.14159265358979323846264338327950288;
constexpr double INV_4PI = 0.07957747154594767;  // 1/(4 pi)

// column classes decided by the prep kernel
enum : uint8_t { COL_FLAT_TRI = 0, COL_FLAT_QUAD = 1, COL_SPECIAL = 2, COL_NONE = 3 };

constexpr int TILE = 128;      // field elements (matrix columns) per far-field tile
constexpr int NQ_TRI = 13;     // TR13 rule  (gauss.rs:84-87; order is always GAU_MIN=4)
constexpr int NQ_QUAD = 16;    // 4x4 Gauss-Legendre (gauss.rs:94-105)
constexpr int NQ_MAX = 16;
constexpr int FAR_NCONST = 16; // per-column constants, see FarConst

// per-column constant slots in far_c[tile][slot][TILE]:
//   n_y (unit normal), J/(4 pi), y_0 (first quadrature point), 9*area*(1+1e-9),
//   e1 = dx/ds, e2 = dx/dt  (y_q = y_0 + a_q e1 + b_q e2 on a flat element)
enum FarConst {
    FC_NX = 0, FC_NY = 1, FC_NZ = 2, FC_J4PI = 3, FC_Y0X = 4, FC_Y0Y = 5, FC_Y0Z = 6, FC_THR = 7,
    FC_E1X = 8, FC_E1Y = 9, FC_E1Z = 10, FC_E2X = 11, FC_E2Y = 12, FC_E2Z = 13, FC_KC = 14, FC_SPARE1 = 15
};

struct cplx {
    double re, im;
};
__host__ __device__ inline cplx C(double re, double im) { return cplx{re, im}; }
__host__ __device__ inline cplx operator+(cplx a, cplx b) { return C(a.re + b.re, a.im + b.im); }
__host__ __device__ inline cplx operator-(cplx a, cplx b) { return C(a.re - b.re, a.im - b.im); }
__host__ __device__ inline cplx operator-(cplx a) { return C(-a.re, -a.im); }
__host__ __device__ inline cplx operator*(cplx a, cplx b) { return C(a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re); }
__host__ __device__ inline cplx operator*(cplx a, double s) { return C(a.re * s, a.im * s); }
__host__ __device__ inline cplx operator/(cplx a, double s) { return C(a.re / s, a.im / s); }
__host__ __device__ inline cplx conj(cplx a) { return C(a.re, -a.im); }
__host__ __device__ inline double norm_sqr(cplx a) { return a.re * a.re + a.im * a.im; }
__host__ __device__ inline cplx& operator+=(cplx& a, cplx b) { a.re += b.re; a.im += b.im; return a; }
__host__ __device__ inline cplx& operator-=(cplx& a, cplx b) { a.re -= b.re; a.im -= b.im; return a; }

// -----------------------------------------------------------------------------
// sincos for the far-field kernel: Cody-Waite reduction by pi/2 (two FMAs; exact
// enough for |x| < 1e5, far above any k*diameter this code meets) + the classic
// degree-13/14 minimax kernels on [-pi/4, pi/4] (fdlibm k_sin / k_cos constants).
// 20 DP-pipe instructions; max abs error ~1.2e-16 on the reduced argument.
// -----------------------------------------------------------------------------
__device__ __forceinline__ void fast_sincos(double x, double& s_out, double& c_out) {
    const double MAGIC = 6755399441055744.0;  // 1.5 * 2^52: rint() via add/sub
    double t = fma(x, 0.6366197723675814, MAGIC);
    int n = __double2loint(t);
    double q = t - MAGIC;
    double r = fma(-q, 1.5707963267948966, x);
    r = fma(-q, 6.123233995736766e-17, r);
    double z = r * r;
    double ps = fma(z, 1.58969099521155010221e-10, -2.50507602534068634195e-08);
    double pc = fma(z, -1.13596475577881948265e-11, 2.08757232129817482790e-09);
    ps = fma(z, ps, 2.75573137070700676789e-06);
    pc = fma(z, pc, -2.75573143513906633035e-07);
    ps = fma(z, ps, -1.98412698298579493134e-04);
    pc = fma(z, pc, 2.48015872894767294178e-05);
    ps = fma(z, ps, 8.33333333332248946124e-03);
    pc = fma(z, pc, -1.38888888888741095749e-03);
    ps = fma(z, ps, -1.66666666666666324348e-01);
    pc = fma(z, pc, 4.16666666666666019037e-02);
    double rz = r * z;
    double zz = z * z;
    double s = fma(rz, ps, r);
    double c = fma(z, -0.5, 1.0);
    c = fma(zz, pc, c);
    // quadrant: x = r + n*pi/2
    double ss = (n & 1) ? c : s;
    double cc = (n & 1) ? s : c;
    if (n & 2) ss = -ss;
    if ((n + 1) & 2) cc = -cc;
    s_out = ss;
    c_out = cc;
}

// Table variant for the far kernel: e^{i kappa r} straight from the distance r.  kappa (= harmonic_factor * k) is folded into
// the reduction constants and the Taylor coefficients, so neither the product kappa*r nor a second Cody-Waite term is computed:
//     q = rint(r * c1),  c1 = 1024 kappa / pi;      t' = r - q * c2,  c2 = pi / (1024 kappa)      (theta = q pi/1024 + kappa t')
//     sin(kappa t') = t' (kappa - kappa^3/6 t'^2),   cos(kappa t') = 1 - kappa^2/2 t'^2 + kappa^4/24 t'^4      (|kappa t'| <= pi/2048:
//     truncation 7e-17 resp. 2e-20), then one complex rotation by the table entry (cos, sin)(q pi/1024): 12 DP-pipe instructions,
//     no quadrant selects.  The only error that grows with the argument is the rounding of c2 (relative 1.1e-16, i.e. at most
//     1.1e-16 * |kappa r| in the phase) -- the same size as the rounding of the product kappa*r that the reference's sin(k*r) starts
//     from.  Checked on the device against a double-double reference by bemb200_selftest_math.
constexpr int SINCOS_TAB = 2048;
constexpr double SINCOS_STEPS_PER_PI = 1024.0;
struct CisConst {
    double c1, c2;       // reduction
    double s1, s3;       // kappa, -kappa^3/6
    double k2h, k4;      // -kappa^2/2, kappa^4/24
};
// host side (the launchers): c2 is rounded ONCE from the long-double quotient
inline CisConst make_cis_const(double kappa) {
    CisConst c;
    c.c1 = kappa * 325.94932345220167;  // 1024/pi
    c.c2 = kappa != 0.0 ? (double)(0.0030679615757712823008853099L / (long double)kappa) : 0.0;  // pi/1024  (kappa = 0: q = 0, sin = 0, cos = 1)
    c.s1 = kappa;
    c.s3 = -(kappa * kappa * kappa) / 6.0;
    c.k2h = -0.5 * kappa * kappa;
    c.k4 = (kappa * kappa) * (kappa * kappa) / 24.0;
    return c;
}
__device__ __forceinline__ void fast_cis_tab(double r, const CisConst& cc, const double2* __restrict__ tab, double& s_out, double& c_out) {
    const double MAGIC = 6755399441055744.0;
    const double tq = fma(r, cc.c1, MAGIC);
    const int idx = __double2loint(tq) & (SINCOS_TAB - 1);
    const double q = tq - MAGIC;
    const double t = fma(-q, cc.c2, r);
    const double z = t * t;
    const double sn = t * fma(z, cc.s3, cc.s1);
    const double cs = fma(z, fma(z, cc.k4, cc.k2h), 1.0);
    const double2 CS = tab[idx];
    c_out = fma(CS.x, cs, -(CS.y * sn));
    s_out = fma(CS.y, cs, CS.x * sn);
}

// 1/sqrt(a) to ~1 ulp: MUFU.RSQ64H seed + one cubically convergent step (5 DP ops)
__device__ __forceinline__ double fast_rsqrt(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    double h = y * y;
    double e = fma(-a, h, 1.0);
    double t = fma(0.375, e, 0.5);
    double ye = y * e;
    return fma(ye, t, y);
}

}  // namespace bemb
