// fp64_operand_probe.cu -- how fast does the FP64 pipe of one B200 issue DFMA as a function of where the operands come from?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_operand_probe tools/fp64_operand_probe.cu && ./fp64_operand_probe
// Every kernel runs ITERS x UNROLL DFMAs per thread on 148 x 4 blocks of 256 threads (8 or 16 warps per scheduler's worth of
// independent chains); reported: TFLOP/s and the fraction of 148 x 64 x 2 x 1.965 GHz.
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 4096

// 1 register source: x = fma(x, a, b), a and b uniform (kernel parameters)
__global__ void __launch_bounds__(256) k_src1(double* out, double a, double b) {
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}
// 2 register sources: x = fma(x, y, b), y per thread
__global__ void __launch_bounds__(256) k_src2(double* out, double a, double b) {
    double x[16], y[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x + i; y[i] = a + 1e-9 * (threadIdx.x + i); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(x[i], y[i], b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}
// 3 distinct register sources: x = fma(y, z, x)
__global__ void __launch_bounds__(256) k_src3(double* out, double a, double b) {
    double x[16], y[16], z[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = threadIdx.x + i; y[i] = a + 1e-9 * (threadIdx.x + i); z[i] = b + 1e-9 * (threadIdx.x * 3 + i); }
    for (int it = 0; it < ITERS; ++it) {
#pragma unroll
        for (int i = 0; i < 16; ++i) x[i] = fma(y[i], z[i], x[i]);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456) out[0] = s;
}
// GEMM-like register tile R x C: acc[i][j] = fma(a[i], x[j], acc[i][j]); a and x change every iteration (cheap integer-free
// update so that the compiler keeps them in registers), j inner: a[i] can sit in the operand reuse cache
template <int R, int C>
__global__ void __launch_bounds__(256) k_tile(double* out, double a0, double b0) {
    double acc[R][C], a[R], x[C];
#pragma unroll
    for (int i = 0; i < R; ++i) a[i] = a0 + 1e-9 * (threadIdx.x + i);
#pragma unroll
    for (int j = 0; j < C; ++j) x[j] = b0 + 1e-9 * (threadIdx.x * 5 + j);
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) acc[i][j] = 0.0;
    for (int it = 0; it < ITERS * 16 / (R * C); ++it) {
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int j = 0; j < C; ++j) acc[i][j] = fma(a[i], x[j], acc[i][j]);
        // rotate the operands (register moves only every R*C DFMAs)
        double t = a[0];
#pragma unroll
        for (int i = 0; i + 1 < R; ++i) a[i] = a[i + 1];
        a[R - 1] = x[0];
#pragma unroll
        for (int j = 0; j + 1 < C; ++j) x[j] = x[j + 1];
        x[C - 1] = t;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) s += acc[i][j];
    if (s == 123.456) out[0] = s;
}

// same tile, boustrophedon order: consecutive DFMAs share a_i along a row and x_j at the turn, so that one of the two
// non-accumulator sources can always come from the operand reuse cache
template <int R, int C>
__global__ void __launch_bounds__(256) k_tile_zz(double* out, double a0, double b0) {
    double acc[R][C], a[R], x[C];
#pragma unroll
    for (int i = 0; i < R; ++i) a[i] = a0 + 1e-9 * (threadIdx.x + i);
#pragma unroll
    for (int j = 0; j < C; ++j) x[j] = b0 + 1e-9 * (threadIdx.x * 5 + j);
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) acc[i][j] = 0.0;
    for (int it = 0; it < ITERS * 16 / (R * C); ++it) {
#pragma unroll
        for (int i = 0; i < R; ++i)
#pragma unroll
            for (int jj = 0; jj < C; ++jj) {
                const int j = (i & 1) ? C - 1 - jj : jj;
                acc[i][j] = fma(a[i], x[j], acc[i][j]);
            }
        double t = a[0];
#pragma unroll
        for (int i = 0; i + 1 < R; ++i) a[i] = a[i + 1];
        a[R - 1] = x[0];
#pragma unroll
        for (int j = 0; j + 1 < C; ++j) x[j] = x[j + 1];
        x[C - 1] = t;
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < R; ++i)
#pragma unroll
        for (int j = 0; j < C; ++j) s += acc[i][j];
    if (s == 123.456) out[0] = s;
}

// FP64 tensor-core shapes (mma.sync ... f64), register-resident, NACC independent accumulator tiles per warp
template <int SHAPE, int NACC>
__global__ void __launch_bounds__(256) k_dmma(double* out, double a0, double b0, int iters) {
    double a[8], b[4], c[NACC][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = a0 + 1e-9 * (threadIdx.x + i);
#pragma unroll
    for (int i = 0; i < 4; ++i) b[i] = b0 + 1e-9 * (threadIdx.x * 3 + i);
#pragma unroll
    for (int n = 0; n < NACC; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) c[n][i] = 0.0;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int n = 0; n < NACC; ++n) {
            if (SHAPE == 884)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(c[n][0]), "+d"(c[n][1]) : "d"(a[n & 7]), "d"(b[n & 3]));
            if (SHAPE == 1684)
                asm volatile("mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};"
                             : "+d"(c[n][0]), "+d"(c[n][1]), "+d"(c[n][2]), "+d"(c[n][3]) : "d"(a[0]), "d"(a[1]), "d"(b[n & 3]));
            if (SHAPE == 1688)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+d"(c[n][0]), "+d"(c[n][1]), "+d"(c[n][2]), "+d"(c[n][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[n & 1]), "d"(b[2 + (n & 1)]));
            if (SHAPE == 16816)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};"
                             : "+d"(c[n][0]), "+d"(c[n][1]), "+d"(c[n][2]), "+d"(c[n][3])
                             : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                               "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
        }
    }
    double s = 0;
#pragma unroll
    for (int n = 0; n < NACC; ++n)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[n][i];
    if (s == 123.456) out[0] = s;
}

template <typename F>
static void run(const char* name, F launch, double dfma_per_thread, int blocks) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch();
    cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; ++r) {
        cudaEventRecord(e0);
        launch();
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double flop = 2.0 * dfma_per_thread * 256.0 * blocks;
    const double tf = flop / (best * 1e-3) / 1e12;
    printf("%-44s %8.3f ms  %6.2f TFLOP/s  %.3f of nominal 37.22\n", name, best, tf, tf / 37.22496);
}

int main() {
    double* out;
    cudaMalloc(&out, 8);
    for (int bps : {4, 8}) {
        const int blocks = 148 * bps;
        printf("---- %d blocks of 256 threads per SM (%d warps per scheduler)\n", bps, bps * 2);
        run("1 register source  x = fma(x, ua, ub)", [&] { k_src1<<<blocks, 256>>>(out, 1.0000001, 1e-9); }, 16.0 * ITERS, blocks);
        run("2 register sources x = fma(x, y, ub)", [&] { k_src2<<<blocks, 256>>>(out, 1.0000001, 1e-9); }, 16.0 * ITERS, blocks);
        run("3 register sources x = fma(y, z, x)", [&] { k_src3<<<blocks, 256>>>(out, 1.0000001, 1e-9); }, 16.0 * ITERS, blocks);
        run("tile 4x4  acc = fma(a_i, x_j, acc)", [&] { k_tile<4, 4><<<blocks, 256>>>(out, 1.0000001, 1e-9); }, 16.0 * ITERS, blocks);
        run("tile 4x8", [&] { k_tile<4, 8><<<blocks, 256>>>(out, 1.0000001, 1e-9); }, 16.0 * ITERS, blocks);
        run("tile 8x8", [&] { k_tile<8, 8><<<blocks, 256>>>(out, 1.0000001, 1e-9); }, 16.0 * ITERS, blocks);
        run("tile 4x4 zigzag", [&] { k_tile_zz<4, 4><<<blocks, 256>>>(out, 1.0000001, 1e-9); }, 16.0 * ITERS, blocks);
        run("tile 8x8 zigzag", [&] { k_tile_zz<8, 8><<<blocks, 256>>>(out, 1.0000001, 1e-9); }, 16.0 * ITERS, blocks);
        {
            const int it = 2048;
            // FMAs per warp instruction: m8n8k4 256, m16n8k4 512, m16n8k8 1024, m16n8k16 2048 -> "DFMA per thread" = / 32
            run("DMMA m8n8k4   x8 accumulators", [&] { k_dmma<884, 8><<<blocks, 256>>>(out, 1.0000001, 1e-9, it); }, 8.0 * it * 256 / 32, blocks);
            run("DMMA m16n8k4  x8", [&] { k_dmma<1684, 8><<<blocks, 256>>>(out, 1.0000001, 1e-9, it); }, 8.0 * it * 512 / 32, blocks);
            run("DMMA m16n8k8  x8", [&] { k_dmma<1688, 8><<<blocks, 256>>>(out, 1.0000001, 1e-9, it); }, 8.0 * it * 1024 / 32, blocks);
            run("DMMA m16n8k16 x8", [&] { k_dmma<16816, 8><<<blocks, 256>>>(out, 1.0000001, 1e-9, it); }, 8.0 * it * 2048 / 32, blocks);
            run("DMMA m16n8k16 x2", [&] { k_dmma<16816, 2><<<blocks, 256>>>(out, 1.0000001, 1e-9, it); }, 2.0 * it * 2048 / 32, blocks);
        }
        run("tile 2x16", [&] { k_tile<2, 16><<<blocks, 256>>>(out, 1.0000001, 1e-9); }, 16.0 * ITERS, blocks);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
