#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b, double c0, double c1) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%4,%5};\n"
                 : "=d"(d0), "=d"(d1) : "d"(a), "d"(b), "d"(c0), "d"(c1));
}
template <int NACC>
__global__ void k_dmma(double* out, int iters) {
    double c[NACC][2];
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x; c[i][1] = i; }
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c[i][0], c[i][1], a, b, c[i][0], c[i][1]);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) printf("  [dmma NACC=%d warps/blk=%d] cycles per DMMA per warp: %.2f\n", NACC, blockDim.x / 32, double(t1 - t0) / (double(iters) * NACC));
}
template <int NACC>
__global__ void k_dfma(double* out, int iters) {
    double c[NACC];
    for (int i = 0; i < NACC; i++) c[i] = threadIdx.x + i;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1e-9 * threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) c[i] = fma(c[i], a, b);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < NACC; i++) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (blockIdx.x == 0 && threadIdx.x == 0) printf("  [dfma NACC=%d warps/blk=%d] cycles per DFMA per warp: %.2f\n", NACC, blockDim.x / 32, double(t1 - t0) / (double(iters) * NACC));
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
    double* out; cudaMalloc(&out, 148 * 1024 * 8 * 8);
    const int iters = 20000;
    for (int wpb : {1, 4, 8, 16, 32}) {
        float ms = timeit([&] { k_dmma<12><<<148, 32 * wpb>>>(out, iters); });
        double fl = 148.0 * wpb * iters * 12 * 512;
        printf("DMMA m8n8k4 x12 acc, %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", wpb, ms, fl / ms * 1e-9);
        ms = timeit([&] { k_dmma<1><<<148, 32 * wpb>>>(out, iters); });
        fl = 148.0 * wpb * iters * 1 * 512;
        printf("DMMA m8n8k4 x1 acc (dependent), %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", wpb, ms, fl / ms * 1e-9);
        ms = timeit([&] { k_dfma<12><<<148, 32 * wpb>>>(out, iters); });
        fl = 148.0 * wpb * iters * 12 * 64;
        printf("DFMA x12, %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", wpb, ms, fl / ms * 1e-9);
        ms = timeit([&] { k_dfma<1><<<148, 32 * wpb>>>(out, iters); });
        fl = 148.0 * wpb * iters * 1 * 64;
        printf("DFMA x1 (dependent), %2d warps/SM: %.3f ms  %.2f TFLOP/s\n", wpb, ms, fl / ms * 1e-9);
    }
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
