// FP64 roofline denominators for the B200 box: cuBLAS DGEMM, raw DMMA (mma.sync m8n8k4 f64) and raw DFMA.
// MEASURED_PEAKS.json carries no FP64 entry, so bench/roofline numbers cite the output of this tool.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/fp64_peak tools/fp64_peak.cu -lcublas
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include <cublas_v2.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double& d0, double& d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void dmma_loop(double* out, int iters, double a0, double b0) {
    double acc[NACC][2];
#pragma unroll
    for (int i = 0; i < NACC; i++) { acc[i][0] = 0.0; acc[i][1] = 0.0; }
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma(acc[i][0], acc[i][1], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i][0] + acc[i][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void dfma_loop(double* out, int iters, double a0, double b0) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = i;
    double a = a0 + threadIdx.x * 1e-9, b = b0;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int nsm = p.multiProcessorCount;
    printf("{\"gpu\": \"%s\", \"sms\": %d", p.name, nsm);
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    // DMMA: warps per SM sweep
    for (int threads : {128, 256, 512, 1024}) {
        for (int cps : {1, 2}) {
            if (threads * cps > 2048) continue;
            int iters = 20000; const int NACC = 16;
            dmma_loop<NACC><<<nsm * cps, threads>>>(out, 100, 1.0, 1.0);
            CK(cudaDeviceSynchronize());
            float best = 1e30f;
            for (int r = 0; r < 3; r++) {
                cudaEventRecord(e0);
                dmma_loop<NACC><<<nsm * cps, threads>>>(out, iters, 1.0, 1.0);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
                cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            double flops = 2.0 * 8 * 8 * 4 * (double)NACC * iters * (threads / 32) * nsm * cps;
            printf(", \"dmma_tflops_t%d_c%d\": %.2f", threads, cps, flops / best * 1e-9);
        }
    }
    for (int threads : {256, 512, 1024}) {
        int iters = 20000; const int NACC = 16;
        dfma_loop<NACC><<<nsm * 2, threads>>>(out, 100, 1.0000001, 1e-9);
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 3; r++) {
            cudaEventRecord(e0);
            dfma_loop<NACC><<<nsm * 2, threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        double flops = 2.0 * (double)NACC * iters * threads * nsm * 2;
        printf(", \"dfma_tflops_t%d\": %.2f", threads, flops / best * 1e-9);
    }
    // cuBLAS DGEMM
    for (int n : {4096, 8192}) {
        double *A, *B, *C; size_t bytes = sizeof(double) * n * n;
        CK(cudaMalloc(&A, bytes)); CK(cudaMalloc(&B, bytes)); CK(cudaMalloc(&C, bytes));
        CK(cudaMemset(A, 0, bytes)); CK(cudaMemset(B, 0, bytes)); CK(cudaMemset(C, 0, bytes));
        cublasHandle_t h; cublasCreate(&h);
        double al = 1.0, be = 0.0;
        for (int r = 0; r < 2; r++) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, A, n, B, n, &be, C, n);
        CK(cudaDeviceSynchronize());
        float best = 1e30f;
        for (int r = 0; r < 10; r++) {
            cudaEventRecord(e0);
            cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, A, n, B, n, &be, C, n);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf(", \"dgemm%d_tflops_burst\": %.2f", n, 2.0 * n * (double)n * n / best * 1e-9);
        if (n == 8192) {
            // sustained: back to back for ~4 s
            int reps = 0; cudaEventRecord(e0);
            float el = 0;
            while (el < 4000.f) {
                for (int r = 0; r < 5; r++) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, A, n, B, n, &be, C, n);
                reps += 5; cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&el, e0, e1);
            }
            printf(", \"dgemm%d_tflops_sustained\": %.2f", n, 2.0 * n * (double)n * n * reps / el * 1e-9);
        }
        // SYRK too (closer to the Cholesky trailing update)
        cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, &al, A, n, &be, C, n);
        CK(cudaDeviceSynchronize());
        best = 1e30f;
        for (int r = 0; r < 5; r++) {
            cudaEventRecord(e0);
            cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, n, n, &al, A, n, &be, C, n);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
            cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf(", \"dsyrk%d_tflops_burst\": %.2f", n, (double)n * n * n / best * 1e-9);
        cublasDestroy(h); cudaFree(A); cudaFree(B); cudaFree(C);
    }
    printf("}\n");
    return 0;
}
