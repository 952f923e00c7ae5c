#include <cstdio>
#include <cuda_runtime.h>
template <int R> __global__ void __launch_bounds__(96) k(float *o) {
    float a[R];
#pragma unroll
    for (int i = 0; i < R; i++) a[i] = o[i + threadIdx.x];
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int i = 0; i < R; i++) a[i] = a[i] * a[(i + 1) % R] + a[(i + 7) % R];
    float s = 0;
#pragma unroll
    for (int i = 0; i < R; i++) s += a[i];
    o[threadIdx.x] = s;
}
template <int R> void test(int threads) {
    cudaFuncAttributes at; cudaFuncGetAttributes(&at, k<R>);
    for (int smem = 8 * 1024; smem <= 24 * 1024; smem += 16 * 1024) {
        int nb = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k<R>, threads, smem);
        printf("R=%d regs=%d threads=%d smem=%d -> blocks/SM=%d\n", R, at.numRegs, threads, smem, nb);
    }
}
int main() { test<56>(96); test<64>(96); test<72>(96); test<80>(96); test<88>(96); test<56>(64); test<88>(64); test<72>(160); test<88>(160); return 0; }
