// Measures the FP64 denominators for the roofline on this box (not product code):
//   - DMMA.8x8x4 issue-bound throughput (registers only)
//   - DFMA throughput
//   - cuBLAS DGEMM 8192^3 / 4096^3 (library reference for the f64 tensor peak)
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 tools/peak_f64.cu -lcublas -o gpurun_out/peak_f64
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

template <int NACC>
__global__ void dmma_rate(double* out, int iters) {
    double c0[NACC], c1[NACC];
    double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
#pragma unroll
    for (int i = 0; i < NACC; ++i) { c0[i] = i; c1[i] = -i; }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_rate(double* out, int iters) {
    double c[NACC];
    double a = 1.0 + threadIdx.x * 1e-9, b = threadIdx.x * 2e-3;
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = i;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < NACC; ++i) c[i] = fma(c[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; ++i) s += c[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    int sms = prop.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * sms * 8 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    printf("{\"gpu\": \"%s\", \"sms\": %d", prop.name, sms);
    for (int warps : {4, 8, 16}) {
        int iters = 20000;
        dmma_rate<16><<<sms, warps * 32>>>(out, 100);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        dmma_rate<16><<<sms, warps * 32>>>(out, iters);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 256 * 16 * (double)iters * warps * sms;
        printf(", \"dmma_tflops_w%d\": %.2f", warps, flops / (ms * 1e-3) / 1e12);
    }
    for (int warps : {8, 16, 32}) {
        int iters = 20000;
        dfma_rate<16><<<sms, warps * 32>>>(out, 100);
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        dfma_rate<16><<<sms, warps * 32>>>(out, iters);
        cudaEventRecord(e1); CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        double flops = 2.0 * 32 * 16 * (double)iters * warps * sms;
        printf(", \"dfma_tflops_w%d\": %.2f", warps, flops / (ms * 1e-3) / 1e12);
    }
    cublasHandle_t h; cublasCreate(&h);
    for (int n : {2048, 4096, 8192}) {
        double *A, *B, *Cm;
        CK(cudaMalloc(&A, sizeof(double) * n * n)); CK(cudaMalloc(&B, sizeof(double) * n * n)); CK(cudaMalloc(&Cm, sizeof(double) * n * n));
        CK(cudaMemset(A, 0, sizeof(double) * n * n)); CK(cudaMemset(B, 0, sizeof(double) * n * n));
        double one = 1.0, zero = 0.0;
        for (int w = 0; w < 2; ++w) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, Cm, n);
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 5; ++r) {
            cudaEventRecord(e0);
            cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_N, n, n, n, &one, A, n, B, n, &zero, Cm, n);
            cudaEventRecord(e1); CK(cudaDeviceSynchronize());
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        printf(", \"cublas_dgemm_%d_tflops\": %.2f", n, 2.0 * n * n * (double)n / (best * 1e-3) / 1e12);
        cudaFree(A); cudaFree(B); cudaFree(Cm);
    }
    printf("}\n");
    return 0;
}
