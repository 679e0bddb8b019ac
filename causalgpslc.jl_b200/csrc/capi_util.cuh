// Helpers for the extern "C" layer: staging of host buffers, argument checks.
#pragma once
#include <vector>
#include "context.cuh"

namespace gpslc {

// Device view of a caller buffer: either the caller's own device pointer, or a temporary device copy of host data.
template <class T>
struct Staged {
    Ctx* ctx; T* d = nullptr; T* host = nullptr; size_t count = 0; bool owned = false; bool out = false;
    Staged(Ctx* c) : ctx(c) {}
    ~Staged() {}   // device copies live in the context's staging arena (ArenaScope of the enclosing call)
    // input buffer
    int in(int loc, const T* p, size_t n) {
        count = n;
        if (!p || n == 0) { d = nullptr; return GPSLC_OK; }
        if (loc == 1) { d = const_cast<T*>(p); return GPSLC_OK; }
        owned = true;
        GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&d), n * sizeof(T)));
        GP_CUDA(ctx, cudaMemcpyAsync(d, p, n * sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
        return GPSLC_OK;
    }
    // output buffer (copied back by finish())
    int outbuf(int loc, T* p, size_t n) {
        count = n; out = true;
        if (!p || n == 0) { d = nullptr; return GPSLC_OK; }
        if (loc == 1) { d = p; return GPSLC_OK; }
        owned = true; host = p;
        GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&d), n * sizeof(T)));
        return GPSLC_OK;
    }
    int finish() {
        if (owned && out && host) GP_CUDA(ctx, cudaMemcpyAsync(host, d, count * sizeof(T), cudaMemcpyDeviceToHost, ctx->stream));
        return GPSLC_OK;
    }
};

#define GP_TRY(expr) do { int _rc = (expr); if (_rc) return _rc; } while (0)

__global__ void inv_sq_kernel(const double* ls, double* w, size_t n, int square);
int launch_inv_sq(Ctx* ctx, const double* ls, double* w, size_t n, int square);

}  // namespace gpslc
