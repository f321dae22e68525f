// direct.cu -- lu_solve (math-solvers/src/direct/lu.rs:136-161), the default solver of
// BemSolver::solve_dense_system (bem_solver.rs:435-441, SolverMethod::Direct).  The reference's
// native build calls LAPACK zgesv through ndarray-linalg; here the same factorisation runs on the
// device through cuSOLVER (a plain library call, like the reference's: zgetrf + zgetrs), bound at
// run time with dlopen so that libbemb200 carries no link-time dependency on it.
//
// The device matrix is row-major, LAPACK is column-major: the buffer read column-major is A^T, so
// A^T = P L U is factored and A x = b is solved with trans = T.
#include <dlfcn.h>

#include <string>

#include "api_internal.h"

using namespace bemb;

namespace {

typedef void* cusolverDnHandle_t;
typedef int cusolverStatus_t;
enum { CUBLAS_OP_N_ = 0, CUBLAS_OP_T_ = 1 };

struct Cusolver {
    void* lib = nullptr;
    cusolverStatus_t (*Create)(cusolverDnHandle_t*) = nullptr;
    cusolverStatus_t (*Destroy)(cusolverDnHandle_t) = nullptr;
    cusolverStatus_t (*SetStream)(cusolverDnHandle_t, cudaStream_t) = nullptr;
    cusolverStatus_t (*ZgetrfBufferSize)(cusolverDnHandle_t, int, int, double2*, int, int*) = nullptr;
    cusolverStatus_t (*Zgetrf)(cusolverDnHandle_t, int, int, double2*, int, double2*, int*, int*) = nullptr;
    cusolverStatus_t (*Zgetrs)(cusolverDnHandle_t, int, int, int, const double2*, int, const int*, double2*, int, int*) = nullptr;
    std::string why;
};

Cusolver& cusolver() {
    static Cusolver cs;
    static bool tried = false;
    if (tried) return cs;
    tried = true;
    const char* names[] = {"libcusolver.so.11", "libcusolver.so", "/usr/local/cuda/lib64/libcusolver.so.11", "libcusolver.so.12"};
    for (const char* nm : names) {
        cs.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (cs.lib) break;
    }
    if (!cs.lib) {
        cs.why = std::string("cannot load libcusolver: ") + (dlerror() ? dlerror() : "?");
        return cs;
    }
#define BIND(field, sym)                                               \
    cs.field = reinterpret_cast<decltype(cs.field)>(dlsym(cs.lib, sym)); \
    if (!cs.field) { cs.why = std::string("libcusolver lacks ") + sym; cs.lib = nullptr; return cs; }
    BIND(Create, "cusolverDnCreate")
    BIND(Destroy, "cusolverDnDestroy")
    BIND(SetStream, "cusolverDnSetStream")
    BIND(ZgetrfBufferSize, "cusolverDnZgetrf_bufferSize")
    BIND(Zgetrf, "cusolverDnZgetrf")
    BIND(Zgetrs, "cusolverDnZgetrs")
#undef BIND
    return cs;
}

}  // namespace

extern "C" int bemb200_lu_solve(const bemb200_matrix* cm, const double* b, double* x_out, int overwrite_matrix, double* factor_ms) {
    bemb200_matrix* m = const_cast<bemb200_matrix*>(cm);
    if (!m || !b || !x_out) return set_error(nullptr, BEMB200_EINVAL, "NULL argument");
    bemb200_ctx* ctx = m->ctx;
    if (m->n_rows != m->n_cols) return set_error(ctx, BEMB200_EINVAL, "lu_solve needs a square matrix");  // LuError::DimensionMismatch
    if (ctx->nranks > 1 || m->r0 != 0 || m->r1 != m->n_rows)
        return set_error(ctx, BEMB200_EUNSUPPORTED, "lu_solve needs the whole matrix on one GPU (use gmres / bicgstab on row-sharded systems)");
    if (m->n_rows > 0x7fffffffull) return set_error(ctx, BEMB200_EINVAL, "matrix too large for the 32-bit LAPACK interface");
    Cusolver& cs = cusolver();
    if (!cs.lib) return set_error(ctx, BEMB200_EUNSUPPORTED, cs.why);
    std::lock_guard<std::mutex> lk(ctx->mu);
    BEMB_CUDA(ctx, cudaSetDevice(ctx->device));
    const int n = (int)m->n_rows;
    cudaStream_t s = ctx->stream;
    cusolverDnHandle_t h = nullptr;
    if (cs.Create(&h) != 0) return set_error(ctx, BEMB200_ECUDA, "cusolverDnCreate failed");
    double2 *lu = nullptr, *work = nullptr, *rhs = nullptr;
    int *ipiv = nullptr, *dinfo = nullptr;
    int lwork = 0, hinfo[2] = {0, 0};
    int rc = BEMB200_OK;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    auto fail = [&](int code, const std::string& msg) { rc = set_error(ctx, code, msg); };
    do {
        if (cs.SetStream(h, s) != 0) { fail(BEMB200_ECUDA, "cusolverDnSetStream failed"); break; }
        if (overwrite_matrix) {
            lu = reinterpret_cast<double2*>(m->A);
        } else {
            // the reference's lu_solve(&a, &b) leaves `a` intact: factor a copy (stream-ordered allocation)
            if (cudaMallocAsync((void**)&lu, (size_t)n * n * sizeof(double2), s) != cudaSuccess) { cudaGetLastError(); fail(BEMB200_ENOMEM, "no memory for the LU copy"); lu = nullptr; break; }
            if (cudaMemcpyAsync(lu, m->A, (size_t)n * n * sizeof(double2), cudaMemcpyDeviceToDevice, s) != cudaSuccess) { fail(BEMB200_ECUDA, "matrix copy failed"); break; }
        }
        if (cs.ZgetrfBufferSize(h, n, n, lu, n, &lwork) != 0) { fail(BEMB200_ECUDA, "cusolverDnZgetrf_bufferSize failed"); break; }
        if (cudaMallocAsync((void**)&work, (size_t)(lwork > 0 ? lwork : 1) * sizeof(double2), s) != cudaSuccess ||
            cudaMallocAsync((void**)&ipiv, (size_t)n * sizeof(int), s) != cudaSuccess ||
            cudaMallocAsync((void**)&dinfo, 2 * sizeof(int), s) != cudaSuccess ||
            cudaMallocAsync((void**)&rhs, (size_t)n * sizeof(double2), s) != cudaSuccess) { cudaGetLastError(); fail(BEMB200_ENOMEM, "no memory for the LU workspace"); break; }
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaMemcpyAsync(rhs, b, (size_t)n * sizeof(double2), cudaMemcpyHostToDevice, s);
        cudaEventRecord(e0, s);
        if (cs.Zgetrf(h, n, n, lu, n, work, ipiv, dinfo) != 0) { fail(BEMB200_ECUDA, "cusolverDnZgetrf failed"); break; }
        cudaEventRecord(e1, s);
        if (cs.Zgetrs(h, CUBLAS_OP_T_, n, 1, lu, n, ipiv, rhs, n, dinfo + 1) != 0) { fail(BEMB200_ECUDA, "cusolverDnZgetrs failed"); break; }
        cudaMemcpyAsync(hinfo, dinfo, 2 * sizeof(int), cudaMemcpyDeviceToHost, s);
        cudaMemcpyAsync(x_out, rhs, (size_t)n * sizeof(double2), cudaMemcpyDeviceToHost, s);
        cudaError_t ce = cudaStreamSynchronize(s);
        if (ce != cudaSuccess) { rc = cuda_fail(ctx, ce, "lu_solve"); break; }
        if (hinfo[0] != 0 || hinfo[1] != 0) { fail(BEMB200_ESINGULAR, "Matrix is singular or nearly singular"); break; }  // LuError::SingularMatrix
        if (factor_ms) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, e0, e1);
            *factor_ms = ms;
        }
    } while (0);
    if (work) cudaFreeAsync(work, s);
    if (ipiv) cudaFreeAsync(ipiv, s);
    if (dinfo) cudaFreeAsync(dinfo, s);
    if (rhs) cudaFreeAsync(rhs, s);
    if (lu && !overwrite_matrix) cudaFreeAsync(lu, s);
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    cudaStreamSynchronize(s);
    cs.Destroy(h);
    return rc;
}
