// probe_lat.cu - dependent-issue latencies (cycles) of the operations on the serial chain of the Jacobi eigen-solve:
// DFMA, DMUL, DADD, FFMA, MUFU.RSQ, F2F conversions, LDS (pointer chase), SHFL, BAR.SYNC with 8 warps.
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__global__ void lat(long long* out, double seed, int* chase) {
    __shared__ int sm[1024];
    const int tid = threadIdx.x;
    for (int i = tid; i < 1024; i += blockDim.x) sm[i] = chase[i];
    __syncthreads();
    long long t0, t1;
    double x = seed, y = seed * 0.5;
    float f = (float)seed;
    // DFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = fma(x, y, y);
    t1 = clock64();
    if (tid == 0) out[0] = (t1 - t0);
    // DMUL
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x * y;
    t1 = clock64();
    if (tid == 0) out[1] = (t1 - t0);
    // DADD
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = x + y;
    t1 = clock64();
    if (tid == 0) out[2] = (t1 - t0);
    // FFMA
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = fmaf(f, 0.999f, 0.001f);
    t1 = clock64();
    if (tid == 0) out[3] = (t1 - t0);
    // MUFU.RSQ
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) f = rsqrtf(f) ;
    t1 = clock64();
    if (tid == 0) out[4] = (t1 - t0);
    // F2F round trip f64 -> f32 -> f64
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = (double)((float)x);
    t1 = clock64();
    if (tid == 0) out[5] = (t1 - t0);
    // LDS pointer chase
    int p = tid & 31;
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) p = sm[p];
    t1 = clock64();
    if (tid == 0) out[6] = (t1 - t0);
    // SHFL chain (64-bit)
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) x = __shfl_sync(0xffffffffu, x, (tid + 1) & 31);
    t1 = clock64();
    if (tid == 0) out[7] = (t1 - t0);
    // BAR.SYNC, all warps of the block
    t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) __syncthreads();
    t1 = clock64();
    if (tid == 0) out[8] = (t1 - t0);
    // DFMA throughput: 8 independent chains per thread
    double a0 = seed, a1 = seed + 1, a2 = seed + 2, a3 = seed + 3, a4 = seed + 4, a5 = seed + 5, a6 = seed + 6, a7 = seed + 7;
    t0 = clock64();
#pragma unroll 4
    for (int i = 0; i < N; ++i) {
        a0 = fma(a0, y, y); a1 = fma(a1, y, y); a2 = fma(a2, y, y); a3 = fma(a3, y, y);
        a4 = fma(a4, y, y); a5 = fma(a5, y, y); a6 = fma(a6, y, y); a7 = fma(a7, y, y);
    }
    t1 = clock64();
    if (tid == 0) out[9] = (t1 - t0);
    if (x + f + p + a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 == 12345.678) out[15] = 1;
}
int main() {
    long long* d; int* c;
    cudaMalloc(&d, 16 * 8); cudaMalloc(&c, 4096);
    int h[1024];
    for (int i = 0; i < 1024; ++i) h[i] = (i * 37 + 11) & 1023;
    cudaMemcpy(c, h, 4096, cudaMemcpyHostToDevice);
    const char* names[] = {"DFMA", "DMUL", "DADD", "FFMA", "MUFU.RSQ(f32)", "F2F f64->f32->f64", "LDS chase", "SHFL.64", "BAR.SYNC", "DFMA x8 indep (per 8)"};
    for (int threads : {32, 256}) {
        lat<<<1, threads>>>(d, 1.0000001, c);
        cudaDeviceSynchronize();
        lat<<<1, threads>>>(d, 1.0000001, c);
        cudaDeviceSynchronize();
        long long o[16];
        cudaMemcpy(o, d, 128, cudaMemcpyDeviceToHost);
        printf("threads=%d:", threads);
        for (int k = 0; k < 10; ++k) printf(" %s=%.1f", names[k], (double)o[k] / N);
        printf("\n");
    }
    return 0;
}
