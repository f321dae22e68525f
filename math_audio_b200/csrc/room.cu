// room.cu -- room-acoustics dense path: point-collocation double-layer
// matrix, incident normal-derivative right-hand side and field evaluation of
//   math-bem/src/room_acoustics/solver.rs
//     element_center_and_normal :38-68      element_area :70-122
//     greens_function_3d :18-24             greens_function_derivative :28-35
//     build_bem_matrix_parallel :448-493    calculate_incident_field_derivative_parallel :638-679
//     calculate_field_pressure_bem_parallel :687-748
//   math-xem-common/src/source.rs  DirectivityPattern::interpolate :59-98, Source::amplitude_towards :203-219
// The matrix lands in an ordinary bemb200_matrix, so DenseOperator / gmres (solve_bem_system,
// solver.rs:412-445) run on it unchanged, row-sharded like the TBEM matrix.
#include <vector>

#include "api_internal.h"

using namespace bemb;

struct bemb200_room_mesh {
    bemb200_ctx* ctx = nullptr;
    uint32_t n = 0;
    double* center = nullptr;  // [n][3]
    double* normal = nullptr;  // [n][3]
    double* area = nullptr;    // [n]
};

namespace {

constexpr double PI_D = 3.14159265358979323846;

struct DevSource {
    double pos[3];
    double amp;            // Source.amplitude * crossover.amplitude_at_frequency(f)
    const double* table;   // [nv][nh] or nullptr (omnidirectional)
    int nh, nv;
};

// element_center_and_normal / element_area (solver.rs:38-122), one thread per element
__global__ void room_geometry_kernel(const double* __restrict__ nodes, const uint32_t* __restrict__ conn, uint32_t n,
                                     double* __restrict__ center, double* __restrict__ normal, double* __restrict__ area) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n) return;
    const uint32_t* c = conn + 4ull * e;
    const int nv = c[3] == 0xFFFFFFFFu ? 3 : 4;
    double p[4][3];
    for (int v = 0; v < nv; ++v)
        for (int d = 0; d < 3; ++d) p[v][d] = nodes[3ull * c[v] + d];
    for (int d = 0; d < 3; ++d) {
        double s = 0.0;  // iterator sum: 0 + n0 + n1 + ...
        for (int v = 0; v < nv; ++v) s = __dadd_rn(s, p[v][d]);
        center[3ull * e + d] = __ddiv_rn(s, (double)nv);
    }
    const double v1[3] = {p[1][0] - p[0][0], p[1][1] - p[0][1], p[1][2] - p[0][2]};
    const double v2[3] = {p[2][0] - p[0][0], p[2][1] - p[0][1], p[2][2] - p[0][2]};
    const double nx = __dsub_rn(__dmul_rn(v1[1], v2[2]), __dmul_rn(v1[2], v2[1]));
    const double ny = __dsub_rn(__dmul_rn(v1[2], v2[0]), __dmul_rn(v1[0], v2[2]));
    const double nz = __dsub_rn(__dmul_rn(v1[0], v2[1]), __dmul_rn(v1[1], v2[0]));
    const double nn = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(nx, nx), __dmul_rn(ny, ny)), __dmul_rn(nz, nz)));
    normal[3ull * e + 0] = __ddiv_rn(nx, nn);
    normal[3ull * e + 1] = __ddiv_rn(ny, nn);
    normal[3ull * e + 2] = __ddiv_rn(nz, nn);
    double a = __dmul_rn(0.5, nn);
    if (nv == 4) {
        const double v3[3] = {p[3][0] - p[0][0], p[3][1] - p[0][1], p[3][2] - p[0][2]};
        const double cx = __dsub_rn(__dmul_rn(v2[1], v3[2]), __dmul_rn(v2[2], v3[1]));
        const double cy = __dsub_rn(__dmul_rn(v2[2], v3[0]), __dmul_rn(v2[0], v3[2]));
        const double cz = __dsub_rn(__dmul_rn(v2[0], v3[1]), __dmul_rn(v2[1], v3[0]));
        a = __dadd_rn(a, __dmul_rn(0.5, __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(cx, cx), __dmul_rn(cy, cy)), __dmul_rn(cz, cz)))));
    }
    area[e] = a;
}

// (ikr - 1) exp(ikr) / (4 pi r^2) * cos_angle     (solver.rs:28-35)
__device__ __forceinline__ cplx dgreen(double r, double k, double cos_angle) {
    if (r < 1e-10) return C(0, 0);
    double sn, cs;
    fast_sincos(k * r, sn, cs);
    const double kr = k * r;
    // (i kr - 1)(cs + i sn) = (-cs - kr sn) + i (kr cs - sn)
    const double sc = cos_angle / (4.0 * PI_D * r * r);
    return C((-cs - kr * sn) * sc, (kr * cs - sn) * sc);
}

// build_bem_matrix_parallel (solver.rs:448-493): A[i,j] = dG/dn(|c_i - c_j|, k, (c_i - c_j).n_i / r) * area_j, diagonal
// (0, -k/(2 pi)) * area_j.  Block = 128 columns x ROWS_PER_BLOCK rows; 512 contiguous bytes per warp and row
// (streaming stores): the kernel is bound by the 16 B/entry it writes.
constexpr int ROOM_TILE = 128;
constexpr int ROOM_ROWS = 64;
__global__ void __launch_bounds__(ROOM_TILE)
room_matrix_kernel(const double* __restrict__ center, const double* __restrict__ normal, const double* __restrict__ area, uint32_t n,
                   double k, uint64_t row_begin, uint64_t row_end, cplx* __restrict__ A) {
    const uint32_t j = blockIdx.x * ROOM_TILE + threadIdx.x;
    const uint64_t r0 = row_begin + (uint64_t)blockIdx.y * ROOM_ROWS;
    const uint64_t r1 = r0 + ROOM_ROWS < row_end ? r0 + ROOM_ROWS : row_end;
    if (j >= n) return;
    const double cjx = center[3ull * j], cjy = center[3ull * j + 1], cjz = center[3ull * j + 2];
    const double aj = area[j];
    for (uint64_t i = r0; i < r1; ++i) {
        cplx v;
        if (i == j) {
            v = C(0.0, -k / (2.0 * PI_D) * aj);
        } else {
            const double dx = __ldg(center + 3 * i) - cjx, dy = __ldg(center + 3 * i + 1) - cjy, dz = __ldg(center + 3 * i + 2) - cjz;
            const double r2 = dx * dx + dy * dy + dz * dz;
            if (r2 < 1e-20) {  // r < 1e-10 (solver.rs:29-31)
                v = C(0, 0);
            } else {
                // 1/r by rsqrt (1 ulp) instead of sqrt + two divisions: the kernel is FP64-issue bound, not store bound
                const double inv_r = fast_rsqrt(r2);
                const double r = r2 * inv_r;
                const double dotn = dx * __ldg(normal + 3 * i) + dy * __ldg(normal + 3 * i + 1) + dz * __ldg(normal + 3 * i + 2);
                const double sc = dotn * inv_r * (inv_r * inv_r) * (aj * (1.0 / (4.0 * PI_D)));  // cos_angle / (4 pi r^2) * area_j
                const double kr = k * r;
                double sn, cs;
                fast_sincos(kr, sn, cs);
                v = C((-cs - kr * sn) * sc, (kr * cs - sn) * sc);
            }
        }
        __stcs(reinterpret_cast<double2*>(A + (i - row_begin) * (uint64_t)n + j), make_double2(v.re, v.im));
    }
}

// DirectivityPattern::interpolate (source.rs:59-98) on a [nv][nh] table sampled every 10 degrees
__device__ double directivity(const DevSource& s, double theta, double phi) {
    if (!s.table) return 1.0;
    const double RAD2DEG = 180.0 / PI_D;
    const double theta_deg = theta * RAD2DEG;
    double phi_deg = phi * RAD2DEG;
    while (phi_deg < 0.0) phi_deg += 360.0;
    while (phi_deg >= 360.0) phi_deg -= 360.0;
    int h_idx = (int)floor(phi_deg / 10.0), v_idx = (int)floor(theta_deg / 10.0);
    h_idx = min(h_idx, s.nh - 1);
    v_idx = min(v_idx, s.nv - 1);
    const int h_next = (h_idx + 1) % s.nh;
    const int v_next = min(v_idx + 1, s.nv - 1);
    const double h_frac = phi_deg / 10.0 - (double)h_idx, v_frac = theta_deg / 10.0 - (double)v_idx;
    const double m00 = s.table[v_idx * s.nh + h_idx], m01 = s.table[v_idx * s.nh + h_next];
    const double m10 = s.table[v_next * s.nh + h_idx], m11 = s.table[v_next * s.nh + h_next];
    const double m0 = m00 * (1.0 - h_frac) + m01 * h_frac;
    const double m1 = m10 * (1.0 - h_frac) + m11 * h_frac;
    return m0 * (1.0 - v_frac) + m1 * v_frac;
}

// Source::amplitude_towards (source.rs:203-219); r = |point - position| already known to be >= 1e-10
__device__ __forceinline__ double amplitude_towards(const DevSource& s, double dx, double dy, double dz, double r) {
    const double theta = acos(dz / r);
    const double phi = atan2(dy, dx);
    return s.amp * directivity(s, theta, phi);
}

// calculate_incident_field_derivative_parallel (solver.rs:638-679): rhs_i = - sum_s dG/dn(r, k, (c_i - x_s).n_i / r) * amp_s(c_i)
__global__ void room_rhs_kernel(const double* __restrict__ center, const double* __restrict__ normal, uint32_t n,
                                const DevSource* __restrict__ sources, int ns, double k, cplx* __restrict__ rhs) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double cx = center[3ull * i], cy = center[3ull * i + 1], cz = center[3ull * i + 2];
    const double nx = normal[3ull * i], ny = normal[3ull * i + 1], nz = normal[3ull * i + 2];
    cplx acc = C(0, 0);
    for (int s = 0; s < ns; ++s) {
        const DevSource src = sources[s];
        const double dx = cx - src.pos[0], dy = cy - src.pos[1], dz = cz - src.pos[2];
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        if (r < 1e-10) continue;
        const double amp = amplitude_towards(src, dx, dy, dz, r);
        const double cosang = (dx * nx + dy * ny + dz * nz) / r;
        const cplx g = dgreen(r, k, cosang);
        acc.re += g.re * amp;
        acc.im += g.im * amp;
    }
    rhs[i] = C(-acc.re, -acc.im);
}

// calculate_field_pressure_bem_parallel (solver.rs:687-748): p(x) = sum_s G(|x - x_s|) amp_s(x)
//   + sum_j dG/dn(|x - c_j|, k, (x - c_j).n_j / r) p_j area_j.   One block per field point.
__global__ void __launch_bounds__(256)
room_field_kernel(const double* __restrict__ center, const double* __restrict__ normal, const double* __restrict__ area, uint32_t n,
                  const DevSource* __restrict__ sources, int ns, const double* __restrict__ pts, const cplx* __restrict__ ps, double k,
                  cplx* __restrict__ out) {
    __shared__ double red[2][8];
    const double x0 = pts[3ull * blockIdx.x], x1 = pts[3ull * blockIdx.x + 1], x2 = pts[3ull * blockIdx.x + 2];
    cplx acc = C(0, 0);
    for (uint32_t j = threadIdx.x; j < n; j += blockDim.x) {
        const double dx = x0 - center[3ull * j], dy = x1 - center[3ull * j + 1], dz = x2 - center[3ull * j + 2];
        const double r = sqrt(dx * dx + dy * dy + dz * dz);
        if (r < 1e-10) continue;
        const double cosang = (dx * normal[3ull * j] + dy * normal[3ull * j + 1] + dz * normal[3ull * j + 2]) / r;
        const cplx g = dgreen(r, k, cosang);
        const cplx t = g * ps[j];
        acc.re += t.re * area[j];
        acc.im += t.im * area[j];
    }
#pragma unroll
    for (int m = 16; m >= 1; m >>= 1) {
        acc.re += __shfl_xor_sync(0xffffffffu, acc.re, m);
        acc.im += __shfl_xor_sync(0xffffffffu, acc.im, m);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = acc.re; red[1][threadIdx.x >> 5] = acc.im; }
    __syncthreads();
    if (threadIdx.x == 0) {
        cplx inc = C(0, 0);
        for (int s = 0; s < ns; ++s) {
            const DevSource src = sources[s];
            const double dx = x0 - src.pos[0], dy = x1 - src.pos[1], dz = x2 - src.pos[2];
            const double r = sqrt(dx * dx + dy * dy + dz * dz);
            if (r < 1e-10) continue;
            const double amp = amplitude_towards(src, dx, dy, dz, r);
            double sn, cs;
            fast_sincos(k * r, sn, cs);
            const double sc = amp / (4.0 * PI_D * r);  // greens_function_3d (solver.rs:18-24)
            inc.re += cs * sc;
            inc.im += sn * sc;
        }
        double tr = 0.0, ti = 0.0;
        for (int w = 0; w < 8; ++w) { tr += red[0][w]; ti += red[1][w]; }
        out[blockIdx.x] = C(inc.re + tr, inc.im + ti);
    }
}

// uploads the sources (and their directivity tables) for one call; everything stream-ordered
struct SourceUpload {
    DevSource* dev = nullptr;
    double* tables = nullptr;
};
int upload_sources(bemb200_ctx* ctx, uint32_t ns, const bemb200_room_source* sources, SourceUpload* up) {
    std::vector<DevSource> hs(ns);
    size_t tab_elems = 0;
    for (uint32_t s = 0; s < ns; ++s)
        if (sources[s].directivity) {
            if (sources[s].n_horizontal == 0 || sources[s].n_vertical == 0 || sources[s].n_horizontal > 360 || sources[s].n_vertical > 181)
                return set_error(ctx, BEMB200_EINVAL, "directivity table needs 1..360 x 1..181 samples");
            tab_elems += (size_t)sources[s].n_horizontal * sources[s].n_vertical;
        }
    cudaStream_t st = ctx->stream;
    BEMB_CUDA(ctx, cudaMallocAsync((void**)&up->dev, ns * sizeof(DevSource), st));
    BEMB_CUDA(ctx, cudaMallocAsync((void**)&up->tables, (tab_elems ? tab_elems : 1) * sizeof(double), st));
    size_t off = 0;
    for (uint32_t s = 0; s < ns; ++s) {
        for (int d = 0; d < 3; ++d) hs[s].pos[d] = sources[s].position[d];
        hs[s].amp = sources[s].amplitude;
        hs[s].table = nullptr;
        hs[s].nh = hs[s].nv = 0;
        if (sources[s].directivity) {
            const size_t cnt = (size_t)sources[s].n_horizontal * sources[s].n_vertical;
            BEMB_CUDA(ctx, cudaMemcpyAsync(up->tables + off, sources[s].directivity, cnt * sizeof(double), cudaMemcpyHostToDevice, st));
            hs[s].table = up->tables + off;
            hs[s].nh = (int)sources[s].n_horizontal;
            hs[s].nv = (int)sources[s].n_vertical;
            off += cnt;
        }
    }
    BEMB_CUDA(ctx, cudaMemcpyAsync(up->dev, hs.data(), ns * sizeof(DevSource), cudaMemcpyHostToDevice, st));
    BEMB_CUDA(ctx, cudaStreamSynchronize(st));  // hs lives on this stack frame
    return BEMB200_OK;
}
void free_sources(bemb200_ctx* ctx, SourceUpload* up) {
    if (up->dev) cudaFreeAsync(up->dev, ctx->stream);
    if (up->tables) cudaFreeAsync(up->tables, ctx->stream);
}

}  // namespace

namespace bemb {
int matrix_alloc_plain(bemb200_ctx* ctx, uint64_t n_rows, uint64_t n_cols, uint64_t r0, uint64_t r1, bemb200_matrix** out);  // api.cu
}

extern "C" {

int bemb200_room_mesh_stage(bemb200_ctx* ctx, const double* nodes, uint64_t n_nodes, const uint32_t* conn, uint64_t n_elem,
                            bemb200_room_mesh** out) {
    if (!ctx) return set_error(nullptr, BEMB200_EINVAL, "ctx is NULL");
    if (!nodes || !conn || !out || n_elem == 0 || n_elem > 0x7fffffffull) return set_error(ctx, BEMB200_EINVAL, "bad argument");
    for (uint64_t e = 0; e < n_elem; ++e)
        for (int v = 0; v < 4; ++v) {
            const uint32_t id = conn[4 * e + v];
            if (id == 0xFFFFFFFFu && v == 3) continue;
            if (id >= n_nodes) return set_error(ctx, BEMB200_EINVAL, "connectivity refers to a node that does not exist");
        }
    std::lock_guard<std::mutex> lk(ctx->mu);
    BEMB_CUDA(ctx, cudaSetDevice(ctx->device));
    bemb200_room_mesh* rm = new bemb200_room_mesh();
    rm->ctx = ctx;
    rm->n = (uint32_t)n_elem;
    double* dnodes = nullptr;
    uint32_t* dconn = nullptr;
    cudaStream_t s = ctx->stream;
    cudaError_t e = cudaMalloc((void**)&rm->center, n_elem * 3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&rm->normal, n_elem * 3 * sizeof(double));
    if (e == cudaSuccess) e = cudaMalloc((void**)&rm->area, n_elem * sizeof(double));
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&dnodes, n_nodes * 3 * sizeof(double), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&dconn, n_elem * 4 * sizeof(uint32_t), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dnodes, nodes, n_nodes * 3 * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dconn, conn, n_elem * 4 * sizeof(uint32_t), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        room_geometry_kernel<<<(unsigned)((n_elem + 127) / 128), 128, 0, s>>>(dnodes, dconn, rm->n, rm->center, rm->normal, rm->area);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (dnodes) cudaFreeAsync(dnodes, s);
    if (dconn) cudaFreeAsync(dconn, s);
    if (e != cudaSuccess) {
        bemb200_room_mesh_free(rm);
        return cuda_fail(ctx, e, "room mesh staging");
    }
    *out = rm;
    return BEMB200_OK;
}

void bemb200_room_mesh_free(bemb200_room_mesh* rm) {
    if (!rm) return;
    cudaSetDevice(rm->ctx->device);
    if (rm->center) cudaFree(rm->center);
    if (rm->normal) cudaFree(rm->normal);
    if (rm->area) cudaFree(rm->area);
    delete rm;
}

uint64_t bemb200_room_mesh_num_elements(const bemb200_room_mesh* rm) { return rm ? rm->n : 0; }

int bemb200_room_mesh_geometry(const bemb200_room_mesh* rm, double* center, double* normal, double* area) {
    if (!rm) return set_error(nullptr, BEMB200_EINVAL, "NULL argument");
    bemb200_ctx* ctx = rm->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    BEMB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (center) BEMB_CUDA(ctx, cudaMemcpyAsync(center, rm->center, (size_t)rm->n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (normal) BEMB_CUDA(ctx, cudaMemcpyAsync(normal, rm->normal, (size_t)rm->n * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (area) BEMB_CUDA(ctx, cudaMemcpyAsync(area, rm->area, (size_t)rm->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    BEMB_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return BEMB200_OK;
}

int bemb200_room_assemble(bemb200_ctx* ctx, const bemb200_room_mesh* rm, double k, uint64_t row_begin, uint64_t row_end,
                          bemb200_matrix** inout, double* kernel_ms) {
    if (!ctx) return set_error(nullptr, BEMB200_EINVAL, "ctx is NULL");
    if (!rm || !inout) return set_error(ctx, BEMB200_EINVAL, "NULL argument");
    if (rm->ctx->device != ctx->device) return set_error(ctx, BEMB200_EINVAL, "room mesh staged on another device");
    if (row_begin > row_end || row_end > rm->n) return set_error(ctx, BEMB200_EINVAL, "row range outside [0, num_elements]");
    if (!(k > 0.0)) return set_error(ctx, BEMB200_EINVAL, "wave number must be > 0");
    std::lock_guard<std::mutex> lk(ctx->mu);
    BEMB_CUDA(ctx, cudaSetDevice(ctx->device));
    bemb200_matrix* m = *inout;
    if (m) {
        if (m->n_rows != rm->n || m->n_cols != rm->n || m->r0 != row_begin || m->r1 != row_end)
            return set_error(ctx, BEMB200_EINVAL, "matrix handle to reuse has a different shape / row range");
    } else {
        int rc = matrix_alloc_plain(ctx, rm->n, rm->n, row_begin, row_end, &m);
        if (rc != BEMB200_OK) return rc;
    }
    cudaStream_t s = ctx->stream;
    const uint64_t nloc = row_end - row_begin;
    cudaError_t e = cudaEventRecord(m->ev[0], s);
    if (e == cudaSuccess && nloc) {
        dim3 grid((rm->n + ROOM_TILE - 1) / ROOM_TILE, (unsigned)((nloc + ROOM_ROWS - 1) / ROOM_ROWS));
        room_matrix_kernel<<<grid, ROOM_TILE, 0, s>>>(rm->center, rm->normal, rm->area, rm->n, k, row_begin, row_end, m->A);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaEventRecord(m->ev[1], s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) {
        if (!*inout) bemb200_matrix_free(m);
        return cuda_fail(ctx, e, "room matrix assembly");
    }
    if (kernel_ms) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, m->ev[0], m->ev[1]);
        *kernel_ms = ms;
    }
    *inout = m;
    return BEMB200_OK;
}

int bemb200_room_incident_rhs(const bemb200_room_mesh* rm, double k, uint32_t n_sources, const bemb200_room_source* sources,
                              double* rhs_host, double* rhs_dev) {
    if (!rm || !sources || (!rhs_host && !rhs_dev)) return set_error(nullptr, BEMB200_EINVAL, "NULL argument");
    bemb200_ctx* ctx = rm->ctx;
    if (n_sources == 0 || n_sources > 4096) return set_error(ctx, BEMB200_EINVAL, "need 1..4096 sources");
    std::lock_guard<std::mutex> lk(ctx->mu);
    BEMB_CUDA(ctx, cudaSetDevice(ctx->device));
    SourceUpload up;
    int rc = upload_sources(ctx, n_sources, sources, &up);
    if (rc != BEMB200_OK) { free_sources(ctx, &up); return rc; }
    cudaStream_t s = ctx->stream;
    cplx* out = reinterpret_cast<cplx*>(rhs_dev);
    cplx* tmp = nullptr;
    cudaError_t e = cudaSuccess;
    if (!out) {
        e = cudaMallocAsync((void**)&tmp, (size_t)rm->n * sizeof(cplx), s);
        out = tmp;
    }
    if (e == cudaSuccess) {
        room_rhs_kernel<<<(rm->n + 127) / 128, 128, 0, s>>>(rm->center, rm->normal, rm->n, up.dev, (int)n_sources, k, out);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess && rhs_host) e = cudaMemcpyAsync(rhs_host, out, (size_t)rm->n * sizeof(cplx), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (tmp) cudaFreeAsync(tmp, s);
    free_sources(ctx, &up);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "room incident rhs");
    return BEMB200_OK;
}

int bemb200_room_field_pressure(const bemb200_room_mesh* rm, double k, uint32_t n_sources, const bemb200_room_source* sources,
                                uint64_t n_points, const double* points, const double* surface_pressure, double* out) {
    if (!rm || !points || !surface_pressure || !out || (n_sources && !sources)) return set_error(nullptr, BEMB200_EINVAL, "NULL argument");
    bemb200_ctx* ctx = rm->ctx;
    if (n_points == 0) return BEMB200_OK;
    if (n_points > 0x7fffffffull || n_sources > 4096) return set_error(ctx, BEMB200_EINVAL, "too many points / sources");
    std::lock_guard<std::mutex> lk(ctx->mu);
    BEMB_CUDA(ctx, cudaSetDevice(ctx->device));
    SourceUpload up;
    if (n_sources) {
        int rc = upload_sources(ctx, n_sources, sources, &up);
        if (rc != BEMB200_OK) { free_sources(ctx, &up); return rc; }
    }
    cudaStream_t s = ctx->stream;
    double* dpts = nullptr;
    cplx *dps = nullptr, *dout = nullptr;
    cudaError_t e = cudaMallocAsync((void**)&dpts, n_points * 3 * sizeof(double), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&dps, (size_t)rm->n * sizeof(cplx), s);
    if (e == cudaSuccess) e = cudaMallocAsync((void**)&dout, n_points * sizeof(cplx), s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dpts, points, n_points * 3 * sizeof(double), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) e = cudaMemcpyAsync(dps, surface_pressure, (size_t)rm->n * sizeof(cplx), cudaMemcpyHostToDevice, s);
    if (e == cudaSuccess) {
        room_field_kernel<<<(unsigned)n_points, 256, 0, s>>>(rm->center, rm->normal, rm->area, rm->n, up.dev, (int)n_sources, dpts, dps, k, dout);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, dout, n_points * sizeof(cplx), cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (dpts) cudaFreeAsync(dpts, s);
    if (dps) cudaFreeAsync(dps, s);
    if (dout) cudaFreeAsync(dout, s);
    free_sources(ctx, &up);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "room field pressure");
    return BEMB200_OK;
}

}  // extern "C"
