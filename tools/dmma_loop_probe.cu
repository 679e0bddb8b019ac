// Development probe: how close can a warp-tile DMMA loop of the factor kernel's shape (16x64 warp tile, m8n8k4, 2 A + 8 B
// fragments per k4 step) get to the raw DMMA issue rate, (V1) with operands in registers, (V2) with the operands loaded from
// shared memory with the kernel's LDS.64 pattern, (V3) with a __syncwarp + mbarrier-like shared-memory poll per slab.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/dmma_loop_probe tools/dmma_loop_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

template <int MODE, int MI, int NI>
__global__ void __launch_bounds__(256, 2) probe(double* out, int slabs, const double* init) {
    extern __shared__ double sh[];   // 3 stages x (2 A slabs + 1 B slab) x 1024 doubles
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 3 * 3 * 1024; i += blockDim.x) sh[i] = init[i % 1024];
    __syncthreads();
    double acc[MI][NI][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
    double ra[MI], rb[NI];
#pragma unroll
    for (int mi = 0; mi < MI; mi++) ra[mi] = init[mi * 32 + lane];
#pragma unroll
    for (int ni = 0; ni < NI; ni++) rb[ni] = init[256 + ni * 32 + lane];
    const int half = warp >> 2, r8base = (warp & 3) * 2;
    volatile int* flag = reinterpret_cast<volatile int*>(sh + 3 * 3 * 1024);
    for (int t = 0; t < slabs; t++) {
        const int st = t % 3;
        const double* sA = sh + st * 3072 + half * 1024 + r8base * 128;
        const double* sB = sh + st * 3072 + 2048;
        if (MODE == 3) { while (flag[st] != 0) {} }
#pragma unroll
        for (int k4 = 0; k4 < 4; k4++) {
            double a[MI], b[NI];
            if (MODE >= 2) {
#pragma unroll
                for (int mi = 0; mi < MI; mi++) a[mi] = sA[mi * 128 + k4 * 32 + lane];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) b[ni] = sB[(ni * 4 + k4) * 32 + lane];
            } else {
#pragma unroll
                for (int mi = 0; mi < MI; mi++) a[mi] = ra[mi];
#pragma unroll
                for (int ni = 0; ni < NI; ni++) b[ni] = rb[ni];
            }
#pragma unroll
            for (int ni = 0; ni < NI; ni++)
#pragma unroll
                for (int mi = 0; mi < MI; mi++) dmma(acc[mi][ni], a[mi], b[ni]);
        }
        if (MODE == 3) { __syncwarp(); if (lane == 0) flag[8 + st] = t; }
    }
    double s = 0;
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < NI; ni++) s += acc[mi][ni][0] + acc[mi][ni][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE, int MI, int NI>
void run(const char* name, int nsm, int cps, double* out, const double* init) {
    const int slabs = 20000;
    const size_t smem = (3 * 3 * 1024 + 64) * sizeof(double);
    CK(cudaFuncSetAttribute(probe<MODE, MI, NI>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    probe<MODE, MI, NI><<<nsm * cps, 256, smem>>>(out, 100, init);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f, ms;
    for (int r = 0; r < 3; r++) {
        cudaEventRecord(e0);
        probe<MODE, MI, NI><<<nsm * cps, 256, smem>>>(out, slabs, init);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1));
        cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
    }
    const double flops = 512.0 * MI * NI * 4 * slabs * 8.0 * nsm * cps;
    printf("%-28s ctas/SM %d: %.2f TFLOP/s\n", name, cps, flops / best * 1e-9);
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int nsm = p.multiProcessorCount;
    double *out, *init;
    CK(cudaMalloc(&out, sizeof(double) * nsm * 2 * 256));
    CK(cudaMalloc(&init, sizeof(double) * 1024));
    CK(cudaMemset(init, 0, sizeof(double) * 1024));
    for (int cps = 1; cps <= 2; cps++) {
        run<1, 2, 8>("V1 regs 16x64", nsm, cps, out, init);
        run<2, 2, 8>("V2 LDS 16x64 (2A+8B)", nsm, cps, out, init);
        run<3, 2, 8>("V3 LDS+poll 16x64", nsm, cps, out, init);
        run<2, 4, 4>("V2 LDS 32x32 (4A+4B)", nsm, cps, out, init);
        run<2, 1, 8>("V2 LDS 8x64 (1A+8B)", nsm, cps, out, init);
        run<2, 2, 4>("V2 LDS 16x32 (2A+4B)", nsm, cps, out, init);
    }
    return 0;
}
