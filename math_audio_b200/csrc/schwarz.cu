// schwarz.cu -- block-Jacobi / additive Schwarz preconditioner built from the assembled dense operator.
//
// Reference semantics: AdditiveSchwarzPreconditioner (math-solvers/src/preconditioners/schwarz.rs).
//   from_csr(matrix, num_subdomains, overlap)   :66-125   contiguous partition (base size n / S, the first n % S subdomains one
//                                                         larger, :71-83), weights 1 / (subdomains containing the DOF) (:92-108)
//   build_subdomain / ilu_factorize             :205-345  local matrix = rows AND columns of the subdomain; ILU(0) keeps the
//                                                         pattern, which for the dense block of a BEM operator is the whole
//                                                         block: LU without pivoting, row by row, l_ik = a_ik * inv(u_kk),
//                                                         pivots with |u_kk| < 1e-30 skipped
//   Subdomain::solve                            :348-380  forward substitution with the unit lower factor, backward
//                                                         substitution, x_i *= inv(u_ii) when |u_ii| > 1e-30
//   apply_sequential                            :399-417  result[g] += solution[l] * weight[g], subdomains in order
// With overlap = 0 this is block-Jacobi on the diagonal blocks A[S_k, S_k] -- the near field of the clusters S_k, the same
// integrals SLFMM's compute_near_block evaluates (math-bem/src/core/assembly/slfmm.rs:538-608), here taken from the TBEM
// matrix that is already in HBM.  Subdomains may also be given explicitly (any index sets, e.g. spatial clusters as in
// slfmm.rs:444-460 build_cluster_dof_mappings; overlapping sets get the reference's weights).
//
// B200 mapping.  Set-up, once per matrix: gather the blocks, factor them in place (one CTA per block, right-looking -- every
// entry receives the same updates in the same order as the reference's row-by-row loop), then run the reference's two
// substitutions on the unit vectors (one thread per column, the factor row broadcast, the solution rows coalesced): the
// result is the explicit inverse block.  Apply, once per Arnoldi step: ONE batched block GEMV (a warp per row, 16-byte
// coalesced loads, r_k staged in shared memory) that streams sum_k |S_k|^2 * 16 bytes -- HBM/L2-bound like the ZGEMV it
// follows, instead of 2 |S_k| dependent substitution steps.  Row-sharded operators: a subdomain must lie inside one
// rank's row block, then M^-1 acts on the rank's slab before the all-gather and needs no communication of its own.
#include "schwarz.h"

#include <algorithm>
#include <atomic>
#include <vector>

using namespace bemb;

namespace bemb {
int nccl_allgather_bytes(bemb200_ctx* ctx, const void* send, void* recv, size_t count_bytes);

namespace {

constexpr int APPLY_ROWS = 32;      // rows of one subdomain handled by one apply CTA
constexpr int APPLY_THREADS = 256;  // 8 warps, 4 rows each
constexpr uint32_t MAX_SUBDOMAIN = 4096;

__device__ __forceinline__ double cnorm(cplx a) { return sqrt(a.re * a.re + a.im * a.im); }  // ComplexField::norm
__device__ __forceinline__ cplx cinv(cplx a) {                                               // num-complex inv(): conj / norm_sqr
    const double ns = a.re * a.re + a.im * a.im;
    return C(a.re / ns, -a.im / ns);
}

// F_k[i][j] = A[S_k[i], S_k[j]]   (build_subdomain, schwarz.rs:223-233)
__global__ void __launch_bounds__(256)
gather_blocks_kernel(const cplx* __restrict__ A, uint64_t lda, uint64_t r0, const uint64_t* __restrict__ sub_off,
                     const uint64_t* __restrict__ inv_off, const uint32_t* __restrict__ idx, cplx* __restrict__ F) {
    const uint32_t k = blockIdx.x;
    const uint64_t o = sub_off[k];
    const uint32_t sz = (uint32_t)(sub_off[k + 1] - o);
    cplx* Fk = F + inv_off[k];
    const uint32_t* id = idx + o;
    for (uint32_t i = blockIdx.y; i < sz; i += gridDim.y) {
        const cplx* row = A + (uint64_t)id[i] * lda + r0;
        for (uint32_t j = threadIdx.x; j < sz; j += blockDim.x) Fk[(uint64_t)i * sz + j] = row[id[j]];
    }
}

// In-place LU without pivoting of every block (ilu_factorize on a dense pattern, schwarz.rs:268-305): for pivot p,
// l_ip = a_ip * inv(u_pp) (i > p), then a_ij -= l_ip * a_pj (i, j > p).  The reference runs the same updates row by row;
// each entry sees them in the same order (increasing p), so the factors agree up to FMA contraction.
__global__ void __launch_bounds__(1024)
lu_nopivot_kernel(const uint64_t* __restrict__ sub_off, const uint64_t* __restrict__ inv_off, cplx* __restrict__ F) {
    const uint32_t k = blockIdx.x;
    const uint32_t sz = (uint32_t)(sub_off[k + 1] - sub_off[k]);
    cplx* Fk = F + inv_off[k];
    __shared__ cplx s_pinv;
    __shared__ int s_skip;
    const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5, nty = blockDim.x >> 5;
    for (uint32_t p = 0; p + 1 < sz; ++p) {
        if (threadIdx.x == 0) {
            const cplx u = Fk[(uint64_t)p * sz + p];
            s_skip = cnorm(u) < 1e-30 ? 1 : 0;  // schwarz.rs:283-285: that pivot eliminates nothing
            s_pinv = s_skip ? C(0, 0) : cinv(u);
        }
        __syncthreads();
        if (!s_skip) {
            const cplx pinv = s_pinv;
            for (uint32_t i = p + 1 + threadIdx.x; i < sz; i += blockDim.x) {
                cplx* a = Fk + (uint64_t)i * sz + p;
                *a = *a * pinv;
            }
            __syncthreads();
            const cplx* prow = Fk + (uint64_t)p * sz;
            for (uint32_t i = p + 1 + ty; i < sz; i += nty) {
                cplx* irow = Fk + (uint64_t)i * sz;
                const cplx l = irow[p];
                // four independent (load, load, fma, store) groups in flight per thread: the step is bound by L2 latency
                uint32_t j = p + 1 + tx;
                for (; j + 96u < sz; j += 128u) {
                    const cplx a0 = irow[j], a1 = irow[j + 32u], a2 = irow[j + 64u], a3 = irow[j + 96u];
                    const cplx b0 = prow[j], b1 = prow[j + 32u], b2 = prow[j + 64u], b3 = prow[j + 96u];
                    irow[j] = a0 - l * b0; irow[j + 32u] = a1 - l * b1; irow[j + 64u] = a2 - l * b2; irow[j + 96u] = a3 - l * b3;
                }
                for (; j < sz; j += 32u) irow[j] = irow[j] - l * prow[j];
            }
        }
        __syncthreads();
    }
}

// X_k = inverse of block k: Subdomain::solve (schwarz.rs:348-380) applied to every unit vector, one thread per column.
// All threads walk the same (i, j): the factor entry is a broadcast load, the solution row X[j][:] a coalesced one.
__global__ void __launch_bounds__(1024)
invert_kernel(const uint64_t* __restrict__ sub_off, const uint64_t* __restrict__ inv_off, const cplx* __restrict__ F,
              cplx* __restrict__ X) {
    const uint32_t k = blockIdx.x;
    const uint32_t sz = (uint32_t)(sub_off[k + 1] - sub_off[k]);
    const cplx* Fk = F + inv_off[k];
    cplx* Xk = X + inv_off[k];
    for (uint32_t c = threadIdx.x; c < sz; c += blockDim.x) {
        // the sums run over four interleaved partial accumulators (j = 4 t + u): four loads in flight per thread instead of a
        // chain of dependent L2 round trips; the reference adds left to right -- a rounding-level difference in the inverse
        for (uint32_t i = 0; i < sz; ++i) {  // forward: y_i = e_i - sum_{j<i} l_ij y_j
            const cplx* Li = Fk + (uint64_t)i * sz;
            cplx p0 = C(0, 0), p1 = C(0, 0), p2 = C(0, 0), p3 = C(0, 0);
            uint32_t j = 0;
            for (; j + 3u < i; j += 4u) {
                const cplx y0 = Xk[(uint64_t)j * sz + c], y1 = Xk[(uint64_t)(j + 1u) * sz + c], y2 = Xk[(uint64_t)(j + 2u) * sz + c],
                           y3 = Xk[(uint64_t)(j + 3u) * sz + c];
                p0 = p0 + Li[j] * y0; p1 = p1 + Li[j + 1u] * y1; p2 = p2 + Li[j + 2u] * y2; p3 = p3 + Li[j + 3u] * y3;
            }
            for (; j < i; ++j) p0 = p0 + Li[j] * Xk[(uint64_t)j * sz + c];
            Xk[(uint64_t)i * sz + c] = C(i == c ? 1.0 : 0.0, 0.0) - ((p0 + p1) + (p2 + p3));
        }
        for (uint32_t i = sz; i-- > 0;) {    // backward: x_i = (y_i - sum_{j>i} u_ij x_j) * inv(u_ii)
            const cplx* Ui = Fk + (uint64_t)i * sz;
            cplx p0 = C(0, 0), p1 = C(0, 0), p2 = C(0, 0), p3 = C(0, 0);
            uint32_t j = i + 1;
            for (; j + 3u < sz; j += 4u) {
                const cplx x0 = Xk[(uint64_t)j * sz + c], x1 = Xk[(uint64_t)(j + 1u) * sz + c], x2 = Xk[(uint64_t)(j + 2u) * sz + c],
                           x3 = Xk[(uint64_t)(j + 3u) * sz + c];
                p0 = p0 + Ui[j] * x0; p1 = p1 + Ui[j + 1u] * x1; p2 = p2 + Ui[j + 2u] * x2; p3 = p3 + Ui[j + 3u] * x3;
            }
            for (; j < sz; ++j) p0 = p0 + Ui[j] * Xk[(uint64_t)j * sz + c];
            cplx acc = Xk[(uint64_t)i * sz + c] - ((p0 + p1) + (p2 + p3));
            const cplx d = Ui[i];
            if (cnorm(d) > 1e-30) acc = acc * cinv(d);
            Xk[(uint64_t)i * sz + c] = acc;
        }
    }
}

// z_k = X_k r_k for a chunk of APPLY_ROWS rows of one subdomain.  Disjoint subdomains: the result goes straight to z
// (weights are 1); overlapping ones park it in `sol` for combine_kernel.
__global__ void __launch_bounds__(APPLY_THREADS)
schwarz_apply_kernel(const uint64_t* __restrict__ sub_off, const uint64_t* __restrict__ inv_off, const uint32_t* __restrict__ idx,
                     const cplx* __restrict__ inv, const uint32_t* __restrict__ cta_sub, const uint32_t* __restrict__ cta_row,
                     const cplx* __restrict__ r_loc, cplx* __restrict__ z_loc, cplx* __restrict__ sol) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* s_r = reinterpret_cast<cplx*>(smem_raw);
    const uint32_t k = cta_sub[blockIdx.x], row0 = cta_row[blockIdx.x];
    const uint64_t o = sub_off[k];
    const uint32_t sz = (uint32_t)(sub_off[k + 1] - o);
    const uint32_t* id = idx + o;
    const cplx* Xk = inv + inv_off[k];
    for (uint32_t c = threadIdx.x; c < sz; c += APPLY_THREADS) s_r[c] = r_loc[id[c]];
    __syncthreads();
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t row_end = row0 + APPLY_ROWS < sz ? row0 + APPLY_ROWS : sz;
    for (uint32_t l = row0 + warp; l < row_end; l += APPLY_THREADS / 32) {
        const double2* row = reinterpret_cast<const double2*>(Xk + (uint64_t)l * sz);
        double ar = 0.0, ai = 0.0;
#pragma unroll 4
        for (uint32_t c = lane; c < sz; c += 32u) {
            const double2 a = row[c];
            const cplx x = s_r[c];
            ar = fma(a.x, x.re, fma(-a.y, x.im, ar));
            ai = fma(a.x, x.im, fma(a.y, x.re, ai));
        }
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) {
            ar += __shfl_xor_sync(0xffffffffu, ar, m);
            ai += __shfl_xor_sync(0xffffffffu, ai, m);
        }
        if (lane == 0) {
            if (sol) sol[o + l] = C(ar, ai);
            else z_loc[id[l]] = C(ar, ai);
        }
    }
}

// The same product for S interleaved right-hand sides (batched GMRES): Z[idx_k[l]][s] = sum_c X_k[l][c] R[idx_k[c]][s].  Lane = right-
// hand side; the inverse-block entry is a warp-uniform load, the R chunk (BLK_CH rows x S) sits in shared memory.  Disjoint
// subdomains only (block-Jacobi).
constexpr int BLK_CH = 128;
__global__ void __launch_bounds__(APPLY_THREADS)
schwarz_apply_block_kernel(const uint64_t* __restrict__ sub_off, const uint64_t* __restrict__ inv_off, const uint32_t* __restrict__ idx,
                           const cplx* __restrict__ inv, const uint32_t* __restrict__ cta_sub, const uint32_t* __restrict__ cta_row,
                           const cplx* __restrict__ R_loc, cplx* __restrict__ Z_loc, int S) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    cplx* s_r = reinterpret_cast<cplx*>(smem_raw);  // [BLK_CH][S]
    const uint32_t k = cta_sub[blockIdx.x], row0 = cta_row[blockIdx.x];
    const uint64_t o = sub_off[k];
    const uint32_t sz = (uint32_t)(sub_off[k + 1] - o);
    const uint32_t* id = idx + o;
    const cplx* Xk = inv + inv_off[k];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    const uint32_t row_end = row0 + APPLY_ROWS < sz ? row0 + APPLY_ROWS : sz;
    constexpr int RPW = APPLY_ROWS / (APPLY_THREADS / 32);  // rows per warp
    double ar[RPW], ai[RPW];
#pragma unroll
    for (int r = 0; r < RPW; ++r) ar[r] = ai[r] = 0.0;
    for (uint32_t c0 = 0; c0 < sz; c0 += BLK_CH) {
        const uint32_t nch = sz - c0 < (uint32_t)BLK_CH ? sz - c0 : (uint32_t)BLK_CH;
        __syncthreads();
        for (uint32_t e = threadIdx.x; e < nch * (uint32_t)S; e += APPLY_THREADS) {
            const uint32_t cc = e / (uint32_t)S, q = e - cc * (uint32_t)S;
            s_r[e] = R_loc[(uint64_t)id[c0 + cc] * S + q];
        }
        __syncthreads();
        if ((int)lane < S) {
#pragma unroll
            for (int r = 0; r < RPW; ++r) {
                const uint32_t l = row0 + warp + (APPLY_THREADS / 32) * r;
                if (l < row_end) {
                    const double2* row = reinterpret_cast<const double2*>(Xk + (uint64_t)l * sz + c0);
                    double xr = ar[r], xi = ai[r];
                    for (uint32_t cc = 0; cc < nch; ++cc) {
                        const double2 a = row[cc];
                        const cplx x = s_r[cc * S + lane];
                        xr = fma(a.x, x.re, fma(-a.y, x.im, xr));
                        xi = fma(a.x, x.im, fma(a.y, x.re, xi));
                    }
                    ar[r] = xr; ai[r] = xi;
                }
            }
        }
    }
    if ((int)lane < S) {
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
            const uint32_t l = row0 + warp + (APPLY_THREADS / 32) * r;
            if (l < row_end) Z_loc[(uint64_t)id[l] * S + lane] = C(ar[r], ai[r]);
        }
    }
}

// result[g] += solution[l] * weight[g], subdomains in their order (schwarz.rs:409-413); rows in no subdomain stay 0
__global__ void __launch_bounds__(256)
schwarz_combine_kernel(uint64_t nloc, const uint64_t* __restrict__ dof_ptr, const uint64_t* __restrict__ dof_pos,
                       const double* __restrict__ weight, const cplx* __restrict__ sol, cplx* __restrict__ z_loc) {
    const uint64_t d = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (d >= nloc) return;
    cplx acc = C(0, 0);
    const double w = weight[d];
    for (uint64_t p = dof_ptr[d]; p < dof_ptr[d + 1]; ++p) acc += sol[dof_pos[p]] * w;
    z_loc[d] = acc;
}

template <typename T>
cudaError_t upload(T** dst, const std::vector<T>& src, cudaStream_t s) {
    *dst = nullptr;
    const size_t bytes = std::max<size_t>(src.size(), 1) * sizeof(T);
    cudaError_t e = cudaMallocAsync((void**)dst, bytes, s);  // stream-ordered, from the cached pool: no device-wide synchronisation
    if (e != cudaSuccess) return e;
    if (!src.empty()) e = cudaMemcpyAsync(*dst, src.data(), src.size() * sizeof(T), cudaMemcpyHostToDevice, s);
    return e;
}

void destroy(bemb200_precond* p) {
    if (!p) return;
    cudaSetDevice(p->ctx->device);
    cudaStream_t s = p->ctx->stream;  // the context outlives its handles (as for matrices)
    void* ptrs[] = {p->sub_off, p->inv_off, p->idx, p->inv, p->sol, p->dof_ptr, p->dof_pos, p->weight, p->cta_sub, p->cta_row, p->tmp};
    for (void* q : ptrs)
        if (q) cudaFreeAsync(q, s);
    delete p;
}

}  // namespace

int schwarz_apply_launches(const bemb200_precond* p) { return p->nsub == 0 ? 0 : (p->disjoint ? 1 : 2); }

cudaError_t schwarz_apply_local(const bemb200_precond* p, const cplx* r_loc, cplx* z_loc, cudaStream_t s) {
    const uint64_t nloc = p->r1 - p->r0;
    if (nloc == 0) return cudaSuccess;
    if (p->nsub == 0) return cudaMemsetAsync(z_loc, 0, nloc * sizeof(cplx), s);
    const size_t smem = (size_t)p->max_size * sizeof(cplx);
    schwarz_apply_kernel<<<p->ncta, APPLY_THREADS, smem, s>>>(p->sub_off, p->inv_off, p->idx, p->inv, p->cta_sub, p->cta_row, r_loc,
                                                              z_loc, p->disjoint ? nullptr : p->sol);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess || p->disjoint) return e;
    schwarz_combine_kernel<<<(unsigned)((nloc + 255) / 256), 256, 0, s>>>(nloc, p->dof_ptr, p->dof_pos, p->weight, p->sol, z_loc);
    return cudaGetLastError();
}

// Z_loc[nloc][S] = M^-1 R_loc[nloc][S] (interleaved right-hand sides); block-Jacobi (disjoint subdomains) only
cudaError_t schwarz_apply_block_local(const bemb200_precond* p, const cplx* R_loc, cplx* Z_loc, int S, cudaStream_t s) {
    const uint64_t nloc = p->r1 - p->r0;
    if (nloc == 0) return cudaSuccess;
    if (!p->disjoint || S < 1 || S > 32) return cudaErrorNotSupported;
    const size_t smem = (size_t)BLK_CH * S * sizeof(cplx);
    static std::atomic<unsigned char> attr_done[64];
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    if (dev < 0 || dev >= 64) dev = 63;
    if (!attr_done[dev].load()) {
        e = cudaFuncSetAttribute(schwarz_apply_block_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(BLK_CH * 32 * sizeof(cplx)));
        if (e != cudaSuccess) return e;
        attr_done[dev].store(1);
    }
    schwarz_apply_block_kernel<<<p->ncta, APPLY_THREADS, smem, s>>>(p->sub_off, p->inv_off, p->idx, p->inv, p->cta_sub, p->cta_row, R_loc,
                                                                    Z_loc, S);
    return cudaGetLastError();
}

}  // namespace bemb

extern "C" {

int bemb200_schwarz_create(const bemb200_matrix* m, uint32_t num_subdomains, const uint64_t* sub_ptr, const uint64_t* sub_idx,
                           bemb200_precond** out) {
    if (!out) return set_error(nullptr, BEMB200_EINVAL, "out is NULL");
    *out = nullptr;
    if (!m) return set_error(nullptr, BEMB200_EINVAL, "matrix is NULL");
    bemb200_ctx* ctx = m->ctx;
    if (m->n_rows != m->n_cols) return set_error(ctx, BEMB200_EINVAL, "a preconditioner needs a square operator");
    if ((sub_ptr == nullptr) != (sub_idx == nullptr)) return set_error(ctx, BEMB200_EINVAL, "sub_ptr and sub_idx go together");
    const uint64_t n = m->n_rows;
    if (n == 0) return set_error(ctx, BEMB200_EINVAL, "empty operator");
    {
        uint64_t b = 0, e = 0;
        bemb200_partition(n, ctx->nranks, ctx->rank, &b, &e);
        if (m->r0 != b || m->r1 != e)
            return set_error(ctx, BEMB200_EINVAL, "matrix slab is not this rank's canonical row block (see bemb200_partition)");
    }
    // ---- the subdomains (global indices), the reference's contiguous partition by default (schwarz.rs:66-83)
    std::vector<uint64_t> ptr, gidx;
    if (!sub_ptr) {
        uint64_t S = num_subdomains < 1 ? 1 : num_subdomains;  // num_subdomains.max(1).min(n)
        if (S > n) S = n;
        const uint64_t base = n / S, rem = n % S;
        ptr.resize(S + 1);
        ptr[0] = 0;
        for (uint64_t i = 0; i < S; ++i) ptr[i + 1] = ptr[i] + base + (i < rem ? 1 : 0);
        gidx.resize(n);
        for (uint64_t i = 0; i < n; ++i) gidx[i] = i;
    } else {
        if (num_subdomains == 0) return set_error(ctx, BEMB200_EINVAL, "no subdomains");
        ptr.assign(sub_ptr, sub_ptr + num_subdomains + 1);
        if (ptr[0] != 0) return set_error(ctx, BEMB200_EINVAL, "sub_ptr[0] must be 0");
        for (uint32_t k = 0; k < num_subdomains; ++k)
            if (ptr[k + 1] < ptr[k]) return set_error(ctx, BEMB200_EINVAL, "sub_ptr must be non-decreasing");
        gidx.assign(sub_idx, sub_idx + ptr[num_subdomains]);
    }
    const uint32_t S = (uint32_t)(ptr.size() - 1);
    const uint64_t chunk = (n + (uint64_t)ctx->nranks - 1) / (uint64_t)ctx->nranks;
    const uint64_t nloc = m->r1 - m->r0;
    std::unique_ptr<bemb200_precond> hp(new bemb200_precond());
    bemb200_precond* p = hp.get();
    p->ctx = ctx; p->n = n; p->r0 = m->r0; p->r1 = m->r1; p->nsub_global = S;
    p->total_entries_global = ptr[S];
    std::vector<uint64_t> sub_off{0}, inv_off{0};
    std::vector<uint32_t> lidx;
    std::vector<uint32_t> count(nloc, 0);
    uint32_t gmin = 0xffffffffu, gmax = 0;
    std::vector<uint32_t> stamp(nloc, 0);  // stamp[li] == k + 1: local row li already occurs in subdomain k
    for (uint32_t k = 0; k < S; ++k) {
        const uint64_t b = ptr[k], e = ptr[k + 1], sz = e - b;
        if (sz == 0) continue;
        if (sz > MAX_SUBDOMAIN) return set_error(ctx, BEMB200_EUNSUPPORTED, "subdomain larger than 4096 unknowns");
        gmin = std::min<uint32_t>(gmin, (uint32_t)sz);
        gmax = std::max<uint32_t>(gmax, (uint32_t)sz);
        uint64_t owner = ~0ull;
        for (uint64_t q = b; q < e; ++q) {
            if (gidx[q] >= n) return set_error(ctx, BEMB200_EINVAL, "subdomain index out of range");
            const uint64_t ow = gidx[q] / chunk;
            if (owner == ~0ull) owner = ow;
            else if (ow != owner)
                return set_error(ctx, BEMB200_EINVAL,
                                 "a subdomain straddles two ranks' row blocks (row-sharded operators need rank-aligned subdomains)");
        }
        if (owner != (uint64_t)ctx->rank) continue;
        for (uint64_t q = b; q < e; ++q) {
            const uint32_t li = (uint32_t)(gidx[q] - m->r0);
            if (stamp[li] == k + 1) return set_error(ctx, BEMB200_EINVAL, "an index occurs twice inside one subdomain");
            stamp[li] = k + 1;
            lidx.push_back(li);
            count[li] += 1;
        }
        sub_off.push_back(lidx.size());
        inv_off.push_back(inv_off.back() + sz * sz);
    }
    p->nsub = (uint32_t)(sub_off.size() - 1);
    p->entries = lidx.size();
    p->inv_elems = inv_off.back();
    p->min_size = gmin == 0xffffffffu ? 0 : gmin;
    p->max_size = gmax;
    p->disjoint = true;
    for (uint64_t d = 0; d < nloc; ++d) p->disjoint = p->disjoint && count[d] == 1;
    uint32_t local_max = 0;
    std::vector<uint32_t> cta_sub, cta_row;
    for (uint32_t k = 0; k < p->nsub; ++k) {
        const uint32_t sz = (uint32_t)(sub_off[k + 1] - sub_off[k]);
        local_max = std::max(local_max, sz);
        for (uint32_t r = 0; r < sz; r += APPLY_ROWS) { cta_sub.push_back(k); cta_row.push_back(r); }
    }
    p->ncta = (uint32_t)cta_sub.size();
    p->max_size = std::max(p->max_size, local_max);

    std::lock_guard<std::mutex> lk(ctx->mu);
    cudaStream_t s = ctx->stream;
#define PC_CUDA(call)                                   \
    do {                                                \
        cudaError_t _e = (call);                        \
        if (_e != cudaSuccess) {                        \
            bemb200_precond* raw = hp.release();        \
            destroy(raw);                               \
            return cuda_fail(ctx, _e, #call);           \
        }                                               \
    } while (0)
    PC_CUDA(cudaSetDevice(ctx->device));
    PC_CUDA(upload(&p->sub_off, sub_off, s));
    PC_CUDA(upload(&p->inv_off, inv_off, s));
    PC_CUDA(upload(&p->idx, lidx, s));
    PC_CUDA(upload(&p->cta_sub, cta_sub, s));
    PC_CUDA(upload(&p->cta_row, cta_row, s));
    PC_CUDA(cudaMallocAsync((void**)&p->tmp, std::max<uint64_t>(nloc, 1) * sizeof(cplx), s));
    if (!p->disjoint) {
        std::vector<uint64_t> dof_ptr(nloc + 1, 0), dof_pos(lidx.size());
        for (uint64_t d = 0; d < nloc; ++d) dof_ptr[d + 1] = dof_ptr[d] + count[d];
        std::vector<uint64_t> fill(dof_ptr.begin(), dof_ptr.end() - 1);
        for (uint64_t q = 0; q < lidx.size(); ++q) dof_pos[fill[lidx[q]]++] = q;  // ascending q = subdomain order
        std::vector<double> w(nloc);
        for (uint64_t d = 0; d < nloc; ++d) w[d] = count[d] ? 1.0 / (double)count[d] : 1.0;
        PC_CUDA(upload(&p->dof_ptr, dof_ptr, s));
        PC_CUDA(upload(&p->dof_pos, dof_pos, s));
        PC_CUDA(upload(&p->weight, w, s));
        PC_CUDA(cudaMallocAsync((void**)&p->sol, std::max<uint64_t>(lidx.size(), 1) * sizeof(cplx), s));
    }
    if (p->max_size * sizeof(cplx) > 48 * 1024)
        PC_CUDA(cudaFuncSetAttribute(schwarz_apply_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(MAX_SUBDOMAIN * sizeof(cplx))));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    if (p->nsub) {
        cplx* F = nullptr;
        PC_CUDA(cudaMallocAsync((void**)&p->inv, p->inv_elems * sizeof(cplx), s));
        cudaError_t fe = cudaMallocAsync((void**)&F, p->inv_elems * sizeof(cplx), s);
        if (fe != cudaSuccess) { PC_CUDA(fe); }
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0, s);
        gather_blocks_kernel<<<dim3(p->nsub, 32), 256, 0, s>>>(m->A, m->n_cols, m->r0, p->sub_off, p->inv_off, p->idx, F);
        const unsigned lu_threads = local_max >= 512 ? 1024u : (local_max >= 128 ? 512u : 256u);
        lu_nopivot_kernel<<<p->nsub, lu_threads, 0, s>>>(p->sub_off, p->inv_off, F);
        unsigned inv_threads = ((local_max + 31u) / 32u) * 32u;
        if (inv_threads > 1024u) inv_threads = 1024u;
        invert_kernel<<<p->nsub, inv_threads, 0, s>>>(p->sub_off, p->inv_off, F, p->inv);
        cudaEventRecord(e1, s);
        cudaError_t ke = cudaGetLastError();
        if (ke == cudaSuccess) ke = cudaStreamSynchronize(s);
        float ms = 0.f;
        if (ke == cudaSuccess) cudaEventElapsedTime(&ms, e0, e1);
        p->factor_ms = ms;
        cudaEventDestroy(e0); cudaEventDestroy(e1);
        cudaFreeAsync(F, s);
        PC_CUDA(ke);
    } else {
        PC_CUDA(cudaStreamSynchronize(s));
    }
#undef PC_CUDA
    *out = hp.release();
    return BEMB200_OK;
}

void bemb200_precond_free(bemb200_precond* p) { destroy(p); }

int bemb200_precond_stats_get(const bemb200_precond* p, bemb200_precond_stats* out) {
    if (!p || !out) return set_error(nullptr, BEMB200_EINVAL, "NULL argument");
    out->num_subdomains = p->nsub_global;
    out->local_subdomains = p->nsub;
    out->min_size = p->min_size;
    out->max_size = p->max_size;
    out->avg_size = p->nsub_global ? (double)p->total_entries_global / (double)p->nsub_global : 0.0;
    out->inverse_bytes = p->inv_elems * sizeof(cplx);
    out->factor_ms = p->factor_ms;
    out->disjoint = p->disjoint ? 1 : 0;
    return BEMB200_OK;
}

// Preconditioner::apply (traits.rs:366-371) with host vectors; collective on a row-sharded operator.
int bemb200_precond_apply(const bemb200_precond* cp, const double* r, double* z) {
    bemb200_precond* p = const_cast<bemb200_precond*>(cp);
    if (!p || !r || !z) return set_error(nullptr, BEMB200_EINVAL, "NULL argument");
    bemb200_ctx* ctx = p->ctx;
    std::lock_guard<std::mutex> lk(ctx->mu);
    BEMB_CUDA(ctx, cudaSetDevice(ctx->device));
    const uint64_t chunk = (p->n + (uint64_t)ctx->nranks - 1) / (uint64_t)ctx->nranks;
    const uint64_t npad = chunk * (uint64_t)ctx->nranks, nloc = p->r1 - p->r0;
    cplx* full = nullptr;
    BEMB_CUDA(ctx, cudaMallocAsync((void**)&full, npad * sizeof(cplx), ctx->stream));
    cudaError_t e = cudaMemsetAsync(full, 0, npad * sizeof(cplx), ctx->stream);
    if (e == cudaSuccess && nloc)
        e = cudaMemcpyAsync(p->tmp, reinterpret_cast<const cplx*>(r) + p->r0, nloc * sizeof(cplx), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = schwarz_apply_local(p, p->tmp, full + p->r0, ctx->stream);
    int rc = BEMB200_OK;
    if (e == cudaSuccess && ctx->nranks > 1) rc = nccl_allgather_bytes(ctx, full + p->r0, full, chunk * sizeof(cplx));
    if (e == cudaSuccess && rc == BEMB200_OK) e = cudaMemcpyAsync(z, full, p->n * sizeof(cplx), cudaMemcpyDeviceToHost, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    cudaFreeAsync(full, ctx->stream);
    if (rc != BEMB200_OK) return rc;
    BEMB_CUDA(ctx, e);
    return BEMB200_OK;
}

}  // extern "C"
