// Dependent-chain latency microbenchmark for the instructions the coder's serial stage is built from.
// One warp, one block; cycles per dependent op.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 --fmad=false
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 2048
#define BENCH(name, init, body)                                              \
    __global__ void k_##name(double *out, long long *cyc, double a0, double b0, int sh) { \
        init;                                                                \
        long long t0 = clock64();                                            \
        _Pragma("unroll 16") for (int i = 0; i < N; i++) { body; }           \
        long long t1 = clock64();                                            \
        if (threadIdx.x == 0) *cyc = t1 - t0;                                \
        out[threadIdx.x] = x + (double)y;                                    \
    }
BENCH(dadd, double x = a0; long long y = 0, x = __dadd_rn(x, b0))
BENCH(dmul, double x = a0; long long y = 0, x = __dmul_rn(x, b0))
BENCH(dfma, double x = a0; long long y = 0, x = __fma_rn(x, b0, b0))
BENCH(ddiv, double x = a0; long long y = 0, x = __ddiv_rn(b0, x))
BENCH(drcp_approx, double x = a0; long long y = 0, asm volatile("rcp.approx.ftz.f64 %0, %1;" : "=d"(x) : "d"(x)))
BENCH(d2ll, double x = a0; long long y = 0, y = __double2ll_rz(x); x = __longlong_as_double(y + 0x4330000000000000LL))
BENCH(d2i, double x = a0; long long y = 0, int q = __double2int_rz(x); x = __hiloint2double(0x43300000, q))
BENCH(ll2d, double x = a0; long long y = 0, x = __ll2double_rn(__double_as_longlong(x) >> 40))
BENCH(i2d, double x = a0; long long y = 0, x = __int2double_rn(__double2hiint(x)))
BENCH(magic_trunc, double x = a0; long long y = 0, double t = __dadd_rn(x, 6755399441055744.0); int q = __double2loint(t); double dr = __dadd_rn(t, -6755399441055744.0); q -= (dr > x); x = __hiloint2double(0x43300000, q))
BENCH(shfl32, double x = a0; long long y = 0; int q = sh, q = __shfl_sync(0xffffffffu, q, (q + 1) & 31); y = q)
BENCH(shfl64, double x = a0; long long y = 0, x = __shfl_sync(0xffffffffu, x, (threadIdx.x + sh) & 31))
BENCH(ballot, double x = a0; long long y = 0; unsigned q = sh, q = __ballot_sync(0xffffffffu, q & (1u << (threadIdx.x & 3))) + threadIdx.x; y = q)
BENCH(iadd, double x = a0; long long y = sh, y = y * 3 + 1)
BENCH(clz, double x = a0; long long y = 0; int q = sh, q = __clz(q) + sh; y = q)
BENCH(shfl_up64_add, double x = a0; long long y = 0, x = __dadd_rn(x, __shfl_up_sync(0xffffffffu, x, 1)))
BENCH(dsetp_sel, double x = a0; long long y = 0, x = (x > b0) ? a0 : b0 + x)

__global__ void k_lds(double *out, long long *cyc, int sh) {
    __shared__ int idx[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) idx[i] = (i + 32 + sh) & 1023;
    __syncwarp();
    int q = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) q = idx[q];
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = t1 - t0;
    out[threadIdx.x] = q;
}
__global__ void k_ldg(double *out, long long *cyc, const int *chain, int mode) {
    int q = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; i++) q = mode ? __ldcg(chain + q) : __ldg(chain + q);
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = t1 - t0;
    out[threadIdx.x] = q;
}
__global__ void k_syncwarp_sts_lds(double *out, long long *cyc, int sh) {
    __shared__ double buf[64];
    double x = threadIdx.x;
    buf[threadIdx.x] = x; __syncwarp();
    long long t0 = clock64();
#pragma unroll 8
    for (int i = 0; i < N; i++) { buf[threadIdx.x] = x; __syncwarp(); x = buf[(threadIdx.x + 1 + sh) & 31]; __syncwarp(); }
    long long t1 = clock64();
    if (threadIdx.x == 0) *cyc = t1 - t0;
    out[threadIdx.x] = x;
}
int main() {
    double *out; long long *cyc; cudaMalloc(&out, 4096); cudaMalloc(&cyc, 8);
    long long h;
#define RUN(name, ...) do { k_##name<<<1, 32>>>(out, cyc, __VA_ARGS__); k_##name<<<1, 32>>>(out, cyc, __VA_ARGS__); cudaDeviceSynchronize(); \
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-18s %7.1f cycles/op\n", #name, (double)h / N); } while (0)
    RUN(dadd, 1.0, 1e-9, 0); RUN(dmul, 1.0, 1.0000001, 0); RUN(dfma, 1.0, 0.5, 0); RUN(ddiv, 1.5, 1.25, 0);
    RUN(drcp_approx, 1.5, 0, 0); RUN(d2ll, 12345.5, 0, 0); RUN(d2i, 12345.5, 0, 0); RUN(ll2d, 1e300, 0, 0); RUN(i2d, 1e300, 0, 0);
    RUN(magic_trunc, 12345.5, 0, 0);
    RUN(shfl32, 0, 0, 1); RUN(shfl64, 1.0, 0, 1); RUN(ballot, 0, 0, 0xff); RUN(iadd, 0, 0, 1); RUN(clz, 0, 0, 5);
    RUN(shfl_up64_add, 1e-9, 0, 0); RUN(dsetp_sel, 1.0, 0.5, 0);
    k_lds<<<1, 32>>>(out, cyc, 0); k_lds<<<1, 32>>>(out, cyc, 0); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-18s %7.1f cycles/op\n", "lds", (double)h / N);
    k_syncwarp_sts_lds<<<1, 32>>>(out, cyc, 0); k_syncwarp_sts_lds<<<1, 32>>>(out, cyc, 0); cudaDeviceSynchronize();
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("%-18s %7.1f cycles/op\n", "sts+sync+lds+sync", (double)h / N);
    // pointer chase: small (L1/L2-resident) and large (DRAM) footprints
    for (int pass = 0; pass < 3; pass++) {
        size_t n = pass == 0 ? (1 << 12) : pass == 1 ? (1 << 22) : (1 << 28);
        int *hc = (int *)malloc(n * 4); int *dc; cudaMalloc(&dc, n * 4);
        size_t stride = pass == 0 ? 33 : 1000003;
        for (size_t i = 0; i < n; i++) hc[i] = (int)((i + stride * 32) % n);
        cudaMemcpy(dc, hc, n * 4, cudaMemcpyHostToDevice);
        for (int mode = 0; mode < 2; mode++) {
            k_ldg<<<1, 32>>>(out, cyc, dc, mode); k_ldg<<<1, 32>>>(out, cyc, dc, mode); cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("ld.%s footprint %zu MiB: %7.1f cycles/op\n", mode ? "cg" : "nc", n * 4 >> 20, (double)h / N);
        }
        cudaFree(dc); free(hc);
    }
    return 0;
}
