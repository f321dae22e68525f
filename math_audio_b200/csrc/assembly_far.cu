// assembly_far.cu -- the FP64 compute-bound far-field assembly kernel (K1).
//
// Replaces, for every (collocation row i, field element j) pair whose level-0
// ratio test passes (dist/sqrt(area) >= 3, singular.rs:553-556), the un-subdivided
// branch of regular_integration (regular.rs:67-154) followed by assemble_tbem for a
// velocity-BC field element (tbem.rs:323-331):
//     A[i,j] = sign*gamma*tau * sum_q H_q  +  beta * sum_q E_q
// with the 13-point triangle rule (Tri3) or the 4x4 Gauss rule (flat Quad4).
//
// Mapping (B200): block = 128 threads <-> one tile of 128 consecutive matrix
// columns; the tile's quadrature points y_q (SoA, 40-48 KB) are brought into
// shared memory by one TMA bulk copy (cp.async.bulk + mbarrier), per-column
// constants (n_y, J/4pi, ratio-test centroid, 9*area, y.n_y) live in registers,
// and the block then walks a chunk of collocation rows: every warp computes 32
// adjacent entries of the same row, so the complex128 stores are 512 contiguous
// bytes per warp.  Source point/normal are block-uniform loads.  Per quadrature
// point the thread spends ~60 DP-pipe instructions (one MUFU-seeded rsqrt, one
// Cody-Waite + minimax sincos, the rest FMAs); nothing else touches HBM.
// Pairs that fail the (guard-banded) ratio test are appended to a compact list
// for the exact near-field kernel.
#include "internal.h"

namespace bemb {

namespace {

__device__ __constant__ double d_wq_tri[NQ_TRI];    // 0.5 * TR13 weights
__device__ __constant__ double d_wq_quad[NQ_QUAD];  // w_i * w_j of the 4x4 rule

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int NQ>
__global__ void __launch_bounds__(TILE, 4)
far_kernel(const double* __restrict__ far_y, const double* __restrict__ far_c, const uint8_t* __restrict__ col_class,
           const double* __restrict__ srcdat, uint32_t n, uint64_t row_begin, uint64_t row_end, uint32_t rows_per_block,
           double wavruim, double k2, double cH, double beta_re, double beta_im, cplx* __restrict__ A, uint64_t lda,
           uint2* __restrict__ near_list, unsigned int near_cap, unsigned int* __restrict__ near_count) {
    extern __shared__ __align__(128) double sm_y[];  // [NQ][3][TILE]
    __shared__ __align__(8) unsigned long long mbar;

    const uint32_t tile = blockIdx.x;
    const uint32_t t = threadIdx.x;
    const uint32_t col = tile * TILE + t;
    constexpr uint32_t BYTES = NQ * 3 * TILE * sizeof(double);
    const uint8_t want = (NQ == NQ_TRI) ? COL_FLAT_TRI : COL_FLAT_QUAD;

    // ---- stage the tile's quadrature points with one TMA bulk copy ------------------
    if (t == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (t == 0) {
        const double* gsrc = far_y + (uint64_t)tile * NQ_MAX * 3 * TILE;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&mbar)), "r"(BYTES) : "memory");
        asm volatile(
            "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sm_y)),
            "l"(gsrc), "r"(BYTES), "r"(smem_u32(&mbar))
            : "memory");
    }
    // per-column constants -> registers (coalesced, overlaps the bulk copy)
    const double* fc = far_c + (uint64_t)tile * FAR_NCONST * TILE + t;
    const double nyx = fc[FC_NX * TILE], nyy = fc[FC_NY * TILE], nyz = fc[FC_NZ * TILE];
    const double j4pi = fc[FC_J4PI * TILE];
    const double ccx = fc[FC_CX * TILE], ccy = fc[FC_CY * TILE], ccz = fc[FC_CZ * TILE];
    const double thr = fc[FC_THR * TILE];
    const double pj = fc[FC_P * TILE];
    const bool active = (col < n) && (col_class[col] == want);
    {
        uint32_t done = 0;
        while (!done) {
            asm volatile(
                "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}\n"
                : "=r"(done)
                : "r"(smem_u32(&mbar)), "r"(0u)
                : "memory");
        }
    }

    const uint64_t r0 = row_begin + (uint64_t)blockIdx.y * rows_per_block;
    const uint64_t r1 = (r0 + rows_per_block < row_end) ? r0 + rows_per_block : row_end;
    const double* wq = (NQ == NQ_TRI) ? d_wq_tri : d_wq_quad;
    const double* yq = sm_y + t;

    for (uint64_t row = r0; row < r1; ++row) {
        const double* sp = srcdat + 8ull * row;  // block-uniform
        const double sx = __ldg(sp + 0), sy = __ldg(sp + 1), sz = __ldg(sp + 2);
        const double nxx = __ldg(sp + 3), nxy = __ldg(sp + 4), nxz = __ldg(sp + 5);
        // level-0 ratio test, guard-banded (exact decision is re-taken by the near kernel)
        const double dcx = ccx - sx, dcy = ccy - sy, dcz = ccz - sz;
        const double d2c = fma(dcz, dcz, fma(dcy, dcy, dcx * dcx));
        const bool is_near = active && (d2c < thr);
        if (__any_sync(0xffffffffu, is_near)) {
            const unsigned mask = __ballot_sync(0xffffffffu, is_near && (uint64_t)col != row);
            if (mask) {
                unsigned base = 0;
                const int leader = __ffs(mask) - 1;
                if ((int)(t & 31) == leader) base = atomicAdd(near_count, __popc(mask));
                base = __shfl_sync(0xffffffffu, base, leader);
                if (is_near && (uint64_t)col != row) {
                    unsigned pos = base + __popc(mask & ((1u << (t & 31)) - 1u));
                    if (pos < near_cap) near_list[pos] = make_uint2((unsigned)row, col);
                }
            }
        }
        // h = (y_q - x).n_y is the same for every point of a flat element
        const double h = pj - fma(sz, nyz, fma(sy, nyy, sx * nyx));
        const double nn = fma(nxz, nyz, fma(nxy, nyy, nxx * nyx));
        double hre = 0.0, him = 0.0, ere = 0.0, eim = 0.0;
#pragma unroll
        for (int q = 0; q < NQ; ++q) {
            const double dx = yq[(q * 3 + 0) * TILE] - sx;
            const double dy = yq[(q * 3 + 1) * TILE] - sy;
            const double dz = yq[(q * 3 + 2) * TILE] - sz;
            const double r2 = fma(dz, dz, fma(dy, dy, dx * dx));
            const double rho = fast_rsqrt(r2);  // 1/r
            const double r = r2 * rho;
            double sn, cs;
            fast_sincos(wavruim * r, sn, cs);
            const double g = (j4pi * rho) * wq[q];  // w J /(4 pi r)
            const double zgr = g * cs, zgi = g * sn;
            const double m = fma(dz, nxz, fma(dy, nxy, dx * nxx));  // (y-x).n_x
            const double re1h = h * rho;                             // u.n_y
            const double re2h = -m * rho;                            // -(u.n_x)
            // zhh_base = zg * (-1/r + i k)
            const double bre = fma(-zgr, rho, -(zgi * wavruim));
            const double bim = fma(zgr, wavruim, -(zgi * rho));
            hre = fma(bre, re1h, hre);
            him = fma(bim, re1h, him);
            const double rq = re1h * re2h;
            const double rho2 = rho * rho;
            const double fre = fma(nn, rho2, fma(3.0, rho2, -k2) * rq);
            const double fim = -(wavruim * rho) * fma(3.0, rq, nn);
            ere = fma(zgr, fre, fma(-zgi, fim, ere));
            eim = fma(zgr, fim, fma(zgi, fre, eim));
        }
        if (active) {
            // coeff = H*sign*gamma*tau + E*beta   (tbem.rs:203,329)
            cplx out;
            out.re = fma(cH, hre, fma(beta_re, ere, -(beta_im * eim)));
            out.im = fma(cH, him, fma(beta_re, eim, beta_im * ere));
            double2 v = make_double2(out.re, out.im);
            __stcs(reinterpret_cast<double2*>(A + (row - row_begin) * lda + col), v);
        }
    }
}

bool g_tables_uploaded[64] = {false};

cudaError_t upload_tables() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 64 && g_tables_uploaded[dev]) return cudaSuccess;
    double wt[NQ_TRI], wqd[NQ_QUAD];
    for (int q = 0; q < NQ_TRI; ++q) wt[q] = hosttab::BEMQ_TR13[q][2] * 0.5;  // gauss.rs:67-89
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) wqd[i * 4 + j] = hosttab::BEMQ_GL4_W[i] * hosttab::BEMQ_GL4_W[j];  // gauss.rs:94-105
    e = cudaMemcpyToSymbol(d_wq_tri, wt, sizeof wt);
    if (e != cudaSuccess) return e;
    e = cudaMemcpyToSymbol(d_wq_quad, wqd, sizeof wqd);
    if (e != cudaSuccess) return e;
    if (dev < 64) g_tables_uploaded[dev] = true;
    return cudaSuccess;
}

template <int NQ>
cudaError_t launch_far_t(const DeviceMesh& m, const Phys& ph, uint64_t row_begin, uint64_t row_end, cplx* A, uint64_t lda,
                         uint2* near_list, unsigned int near_cap, unsigned int* near_count, cudaStream_t s) {
    const uint64_t nrows = row_end - row_begin;
    // enough blocks to fill 148 SMs x 4 resident blocks several times over, but rows
    // chunks long enough to amortise the tile load
    uint32_t rpb = 128;
    while (rpb > 8 && (uint64_t)m.ntiles * ((nrows + rpb - 1) / rpb) < 148ull * 4ull * 4ull) rpb >>= 1;
    dim3 grid(m.ntiles, (unsigned)((nrows + rpb - 1) / rpb));
    const size_t smem = (size_t)NQ * 3 * TILE * sizeof(double);
    static bool attr_set[2] = {false, false};
    const int ai = (NQ == NQ_TRI) ? 0 : 1;
    if (!attr_set[ai]) {
        cudaError_t e = cudaFuncSetAttribute(far_kernel<NQ>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set[ai] = true;
    }
    const double cH = ph.sign * ph.gamma * ph.tau;
    far_kernel<NQ><<<grid, TILE, smem, s>>>(m.far_y, m.far_c, m.col_class, m.src, m.n, row_begin, row_end, rpb, ph.wavruim,
                                            ph.k2, cH, ph.beta.re, ph.beta.im, A, lda, near_list, near_cap, near_count);
    return cudaGetLastError();
}

}  // namespace

int far_kernel_launch_count(const DeviceMesh& m) { return (m.n_flat_tri ? 1 : 0) + (m.n_flat_quad ? 1 : 0); }

cudaError_t launch_far(const DeviceMesh& m, const Phys& ph, uint64_t row_begin, uint64_t row_end, cplx* A, uint64_t lda,
                       uint2* near_list, unsigned int near_cap, unsigned int* near_count, cudaStream_t s) {
    if (row_end <= row_begin) return cudaSuccess;
    cudaError_t e = upload_tables();
    if (e != cudaSuccess) return e;
    if (m.n_flat_tri) {
        e = launch_far_t<NQ_TRI>(m, ph, row_begin, row_end, A, lda, near_list, near_cap, near_count, s);
        if (e != cudaSuccess) return e;
    }
    if (m.n_flat_quad) {
        e = launch_far_t<NQ_QUAD>(m, ph, row_begin, row_end, A, lda, near_list, near_cap, near_count, s);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

}  // namespace bemb
