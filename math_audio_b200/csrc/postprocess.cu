// postprocess.cu -- the O(N) / O(M N) neighbours of the hot path:
//   incident_rhs_kernel      IncidentField::compute_rhs_with_beta   math-bem/src/core/incident.rs:93-342
//   scattered_field_kernel   compute_scattered_field                 math-bem/src/core/postprocess/pressure.rs:81-259
//   rcs_kernel               compute_rcs                             math-bem/src/core/postprocess/pressure.rs:438-478
// All work on the staged (DOF-ordered) mesh so that a frequency sweep never leaves the device.
#include <vector>

#include "api_internal.h"

using namespace bemb;

namespace {

struct Source {
    int kind;       // 0 plane wave (vec = unit direction), 1 point source (vec = position)
    double v[3];
    cplx amp;
};

// rhs_i = -(gamma p_inc + beta tau dp_inc/dn) summed over the sources (incident.rs:317-342)
__global__ void incident_rhs_kernel(const double* __restrict__ src, uint32_t n, const Source* __restrict__ sources, int ns, double k,
                                    double gamma, double tau, cplx beta, cplx* __restrict__ rhs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double* p = src + 8ull * i;
    const double* nr = p + 3;
    cplx pinc = C(0, 0), dpdn = C(0, 0);
    for (int s = 0; s < ns; ++s) {
        const Source sc = sources[s];
        if (sc.kind == 0) {  // incident.rs:103-118, 191-211
            const double kdotx = k * (sc.v[0] * p[0] + sc.v[1] * p[1] + sc.v[2] * p[2]);
            const double kdotn = k * (sc.v[0] * nr[0] + sc.v[1] * nr[1] + sc.v[2] * nr[2]);
            double sn, cs;
            sincos(kdotx, &sn, &cs);
            const cplx pw = sc.amp * C(cs, sn);
            pinc += pw;
            dpdn += C(0.0, kdotn) * pw;
        } else {  // incident.rs:120-134, 213-238
            const double dx = p[0] - sc.v[0], dy = p[1] - sc.v[1], dz = p[2] - sc.v[2];
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            if (r > 1e-10) {
                double sn, cs;
                sincos(k * r, &sn, &cs);
                const cplx g = C(cs, sn) / (4.0 * PI * r);
                pinc += sc.amp * g;
                const cplx dgdr = (C(0.0, k) - C(1.0 / r, 0.0)) * g;
                const double drdn = (dx * nr[0] + dy * nr[1] + dz * nr[2]) / r;
                dpdn += sc.amp * dgdr * drdn;
            }
        }
    }
    const cplx v = -(pinc * gamma + beta * C(tau, 0.0) * dpdn);
    rhs[i] = v;
}

__device__ __constant__ double d_tr7[7][3] = {
    {0.333333333333333, 0.333333333333333, 0.225},
    {0.797426985353087, 0.101286507323456, 0.125939180544827},
    {0.101286507323456, 0.797426985353087, 0.125939180544827},
    {0.101286507323456, 0.101286507323456, 0.125939180544827},
    {0.470142064105115, 0.059715871789770, 0.132394152788506},
    {0.059715871789770, 0.470142064105115, 0.132394152788506},
    {0.470142064105115, 0.470142064105115, 0.132394152788506},
};

// p_scat(x) = sum_j  p_j int dG/dn_y - v_j int G  with the 7-point rule on the element's FIRST
// triangle (Quad4 is approximated by its first three nodes exactly as pressure.rs:184-198 does).
// One block per evaluation point, threads over elements, deterministic block reduction.
__global__ void __launch_bounds__(256)
scattered_field_kernel(const double* __restrict__ coords, uint32_t n, const double* __restrict__ eval, const cplx* __restrict__ ps,
                       const cplx* __restrict__ vs, double wavruim, cplx* __restrict__ out) {
    __shared__ double red[2][8];
    const double x0 = eval[3 * blockIdx.x], x1 = eval[3 * blockIdx.x + 1], x2 = eval[3 * blockIdx.x + 2];
    cplx acc = C(0, 0);
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double* c = coords + 12ull * j;
        const double e1[3] = {c[3] - c[0], c[4] - c[1], c[5] - c[2]};
        const double e2[3] = {c[6] - c[0], c[7] - c[1], c[8] - c[2]};
        const double nv[3] = {e1[1] * e2[2] - e1[2] * e2[1], e1[2] * e2[0] - e1[0] * e2[2], e1[0] * e2[1] - e1[1] * e2[0]};
        const double jac = sqrt(nv[0] * nv[0] + nv[1] * nv[1] + nv[2] * nv[2]);
        if (jac < 1e-15) continue;
        const double en[3] = {nv[0] / jac, nv[1] / jac, nv[2] / jac};
        const cplx p_surf = ps[j];
        const cplx v_surf = vs ? vs[j] : C(0, 0);
        const bool has_v = sqrt(v_surf.re * v_surf.re + v_surf.im * v_surf.im) > 1e-15;
        cplx res = C(0, 0);
#pragma unroll
        for (int q = 0; q < 7; ++q) {
            const double xi = d_tr7[q][0], eta = d_tr7[q][1], wq = d_tr7[q][2] * 0.5;
            const double l0 = 1.0 - xi - eta;
            const double rv[3] = {l0 * c[0] + xi * c[3] + eta * c[6] - x0, l0 * c[1] + xi * c[4] + eta * c[7] - x1,
                                  l0 * c[2] + xi * c[5] + eta * c[8] - x2};
            const double r = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
            if (r < 1e-15) continue;
            const double vjacwe = jac * wq;
            double sn, cs;
            sincos(wavruim * r, &sn, &cs);
            const double re1 = 4.0 * PI * r;
            const cplx zg = C(cs / re1, sn / re1);
            const cplx zgikr = zg * C(-1.0 / r, wavruim);
            const double drdn = (rv[0] * en[0] + rv[1] * en[1] + rv[2] * en[2]) / r;
            res += p_surf * (zgikr * drdn) * vjacwe;
            if (has_v) res -= v_surf * zg * vjacwe;
        }
        acc += res;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        acc.re += __shfl_xor_sync(0xffffffffu, acc.re, m);
        acc.im += __shfl_xor_sync(0xffffffffu, acc.im, m);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc.re; red[1][threadIdx.x >> 5] = acc.im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        cplx t = C(0, 0);
        for (int w = 0; w < 8; ++w) { t.re += red[0][w]; t.im += red[1][w]; }
        out[blockIdx.x] = t;
    }
}

// compute_rcs (pressure.rs:438-478): F(d) = sum_j p_j exp(-i k c_j.d) A_j (i k)(n_j.d), RCS = 4 pi |F|^2.
// One block per direction, threads over elements, deterministic block reduction.
__global__ void __launch_bounds__(256)
rcs_kernel(const double* __restrict__ src, const double* __restrict__ area, uint32_t n, const double* __restrict__ dirs,
           const cplx* __restrict__ ps, double k, double* __restrict__ out) {
    __shared__ double red[2][8];
    const double d0 = dirs[3 * blockIdx.x], d1 = dirs[3 * blockIdx.x + 1], d2 = dirs[3 * blockIdx.x + 2];
    cplx acc = C(0, 0);
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double* c = src + 8ull * j;
        const double phase = -k * (c[0] * d0 + c[1] * d1 + c[2] * d2);
        double sn, cs;
        sincos(phase, &sn, &cs);
        const double n_dot_d = c[3] * d0 + c[4] * d1 + c[5] * d2;
        const cplx t = ((ps[j] * C(cs, sn)) * area[j]) * C(0.0, k);
        acc += t * n_dot_d;
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        acc.re += __shfl_xor_sync(0xffffffffu, acc.re, m);
        acc.im += __shfl_xor_sync(0xffffffffu, acc.im, m);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc.re; red[1][threadIdx.x >> 5] = acc.im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double tr = 0.0, ti = 0.0;
        for (int w = 0; w < 8; ++w) { tr += red[0][w]; ti += red[1][w]; }
        out[blockIdx.x] = 4.0 * 3.14159265358979323846 * (tr * tr + ti * ti);
    }
}

}  // namespace

extern "C" int bemb200_incident_rhs(const bemb200_staged_mesh* sm, const bemb200_physics* phys, double beta_re, double beta_im,
                                    uint32_t n_sources, const int32_t* kinds, const double* vecs, const double* amps,
                                    double* rhs_host, double* rhs_dev) {
    if (!sm || !phys || !kinds || !vecs || !amps || (!rhs_host && !rhs_dev)) return set_error(nullptr, BEMB200_EINVAL, "NULL argument");
    bemb200_ctx* ctx = sm->ctx;
    if (n_sources == 0 || n_sources > 4096) return set_error(ctx, BEMB200_EINVAL, "need 1..4096 sources");
    std::lock_guard<std::mutex> lk(ctx->mu);
    BEMB_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t n = sm->dm.n;
    std::vector<Source> hs(n_sources);
    for (uint32_t s = 0; s < n_sources; ++s) {
        if (kinds[s] != 0 && kinds[s] != 1) return set_error(ctx, BEMB200_EINVAL, "source kind must be 0 (plane wave) or 1 (point source)");
        hs[s].kind = kinds[s];
        for (int d = 0; d < 3; ++d) hs[s].v[d] = vecs[3 * s + d];
        hs[s].amp = C(amps[2 * s], amps[2 * s + 1]);
    }
    Source* dsrc = nullptr;
    cplx* dout = (cplx*)rhs_dev;
    cplx* tmp = nullptr;
    BEMB_CUDA(ctx, cudaMallocAsync((void**)&dsrc, n_sources * sizeof(Source), ctx->stream));
    if (!dout) {
        BEMB_CUDA(ctx, cudaMallocAsync((void**)&tmp, (size_t)n * sizeof(cplx), ctx->stream));
        dout = tmp;
    }
    cudaError_t e = cudaMemcpyAsync(dsrc, hs.data(), n_sources * sizeof(Source), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) {
        incident_rhs_kernel<<<(n + 255) / 256, 256, 0, ctx->stream>>>(sm->dm.src, n, dsrc, (int)n_sources, phys->wave_number, phys->gamma,
                                                                      phys->tau, C(beta_re, beta_im), dout);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && rhs_host) e = cudaMemcpyAsync(rhs_host, dout, (size_t)n * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFreeAsync(dsrc, ctx->stream);
    if (tmp) cudaFreeAsync(tmp, ctx->stream);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "incident_rhs");
    return BEMB200_OK;
}

extern "C" int bemb200_scattered_field(const bemb200_staged_mesh* sm, const bemb200_physics* phys, uint64_t n_eval, const double* eval_pts,
                                       const double* surface_pressure, const double* surface_velocity, double* out) {
    if (!sm || !phys || !eval_pts || !surface_pressure || !out) return set_error(nullptr, BEMB200_EINVAL, "NULL argument");
    bemb200_ctx* ctx = sm->ctx;
    if (n_eval == 0) return BEMB200_OK;
    if (n_eval > 0x7fffffffull) return set_error(ctx, BEMB200_EINVAL, "too many evaluation points");
    std::lock_guard<std::mutex> lk(ctx->mu);
    BEMB_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t n = sm->dm.n;
    double* dev = nullptr;
    cplx *dps = nullptr, *dvs = nullptr, *dout = nullptr;
    cudaStream_t s = ctx->stream;
    BEMB_CUDA(ctx, cudaMallocAsync((void**)&dev, n_eval * 3 * sizeof(double), s));
    BEMB_CUDA(ctx, cudaMallocAsync((void**)&dps, (size_t)n * sizeof(cplx), s));
    BEMB_CUDA(ctx, cudaMallocAsync((void**)&dout, n_eval * sizeof(cplx), s));
    if (surface_velocity) BEMB_CUDA(ctx, cudaMallocAsync((void**)&dvs, (size_t)n * sizeof(cplx), s));
    cudaError_t e = cudaMemcpyAsync(dev, eval_pts, n_eval * 3 * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dps, surface_pressure, (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && dvs) e = cudaMemcpyAsync(dvs, surface_velocity, (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        scattered_field_kernel<<<(unsigned)n_eval, 256, 0, s>>>(sm->dm.coords, n, dev, dps, dvs, phys->wave_number * phys->harmonic_factor, dout);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, n_eval * sizeof(cplx), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFreeAsync(dev, s); cudaFreeAsync(dps, s); cudaFreeAsync(dout, s);
    if (dvs) cudaFreeAsync(dvs, s);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "scattered_field");
    return BEMB200_OK;
}

extern "C" int bemb200_compute_rcs(const bemb200_staged_mesh* sm, const bemb200_physics* phys, uint32_t n_dirs, const double* dirs,
                                   const double* surface_pressure, double* rcs_out) {
    if (!sm || !phys || !dirs || !surface_pressure || !rcs_out) return set_error(nullptr, BEMB200_EINVAL, "NULL argument");
    bemb200_ctx* ctx = sm->ctx;
    if (n_dirs == 0) return BEMB200_OK;
    std::lock_guard<std::mutex> lk(ctx->mu);
    BEMB_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint32_t n = sm->dm.n;
    double *ddirs = nullptr, *dout = nullptr;
    cplx* dps = nullptr;
    cudaStream_t s = ctx->stream;
    BEMB_CUDA(ctx, cudaMallocAsync((void**)&ddirs, (size_t)n_dirs * 3 * sizeof(double), s));
    BEMB_CUDA(ctx, cudaMallocAsync((void**)&dout, (size_t)n_dirs * sizeof(double), s));
    BEMB_CUDA(ctx, cudaMallocAsync((void**)&dps, (size_t)(n ? n : 1) * sizeof(cplx), s));
    cudaError_t e = cudaMemcpyAsync(ddirs, dirs, (size_t)n_dirs * 3 * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess && n) e = cudaMemcpyAsync(dps, surface_pressure, (size_t)n * sizeof(cplx), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        rcs_kernel<<<n_dirs, 256, 0, s>>>(sm->dm.src, sm->dm.area, n, ddirs, dps, phys->wave_number, dout);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(rcs_out, dout, (size_t)n_dirs * sizeof(double), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    cudaFreeAsync(ddirs, s); cudaFreeAsync(dout, s); cudaFreeAsync(dps, s);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "compute_rcs");
    return BEMB200_OK;
}
