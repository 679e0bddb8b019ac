// Library context: one per GPU, owns the stream and the reusable device workspaces. Not thread-safe.
#pragma once
#include <string>
#include <vector>
#include "common.cuh"
#include "factor.cuh"

namespace gpslc {

struct Ctx {
    int device = 0;
    int num_sms = 0;
    cudaStream_t stream = nullptr;
    std::string last_error;
    // factor workspaces: one slot per resident CTA
    int slots = 0;
    size_t slot_scratch_d = 0;   // doubles of L scratch per slot (stride used by the current launch)
    size_t slot_z_d = 0;         // doubles of z/w scratch per slot
    size_t scratch_cap_d = 0, z_cap_d = 0;   // allocated capacities (doubles)
    double* scratch = nullptr;
    double* zbuf = nullptr;
    unsigned int* counter = nullptr;  // work-queue counters (device)
    unsigned long long launches = 0;  // kernels launched by this context (bench.py reports it)
    // Staging arena for host-buffer calls: a bump allocator that is reset at the start of every C-ABI call and grows to the
    // peak demand seen so far (up to ARENA_MAX), so a steady stream of calls does no cudaMalloc / cudaFree at all (cudaFree
    // synchronises the device). Requests that do not fit are served by cudaMalloc and freed when the call returns.
    static constexpr size_t ARENA_MAX = (size_t)1 << 30;
    char* arena = nullptr;
    size_t arena_cap = 0, arena_off = 0, arena_peak = 0;
    std::vector<void*> arena_overflow;
    void arena_begin() {
        if (arena_peak > arena_cap && arena_cap < ARENA_MAX) {
            size_t want = arena_peak < ARENA_MAX ? arena_peak : ARENA_MAX;
            if (arena) cudaFree(arena);
            arena = nullptr; arena_cap = 0;
            if (cudaMalloc(&arena, want) == cudaSuccess) arena_cap = want; else { arena = nullptr; cudaGetLastError(); }
        }
        arena_off = 0; arena_peak = 0;
    }
    void arena_end() {
        for (void* p : arena_overflow) cudaFree(p);
        arena_overflow.clear();
    }
    // Size-keyed cache of long-lived device blocks (chain state, sample buffers of a sampler): a host that creates and destroys
    // samplers of the same shape in a loop (bench.py's e2e leg, SBC studies) re-uses the blocks instead of paying a cudaMalloc
    // and a synchronising cudaFree for each of the ~25 buffers of every sampler. Bounded by BLOCK_CACHE_MAX bytes.
    static constexpr size_t BLOCK_CACHE_MAX = (size_t)4 << 30;
    std::vector<std::pair<size_t, void*>> block_cache;
    size_t block_cache_bytes = 0;
    static size_t block_round(size_t bytes) { return (bytes + 511) & ~(size_t)511; }
    cudaError_t block_alloc(void** p, size_t bytes) {
        bytes = block_round(bytes);
        for (size_t i = 0; i < block_cache.size(); i++)
            if (block_cache[i].first == bytes) {
                *p = block_cache[i].second;
                block_cache_bytes -= bytes;
                block_cache[i] = block_cache.back();
                block_cache.pop_back();
                return cudaSuccess;
            }
        cudaError_t e = cudaMalloc(p, bytes);
        if (e != cudaSuccess && !block_cache.empty()) {      // out of memory with blocks parked in the cache: release them and retry
            cudaGetLastError();
            block_cache_release();
            e = cudaMalloc(p, bytes);
        }
        return e;
    }
    void block_free(void* p, size_t bytes) {
        if (!p) return;
        bytes = block_round(bytes);
        if (block_cache_bytes + bytes <= BLOCK_CACHE_MAX) { block_cache.push_back({bytes, p}); block_cache_bytes += bytes; }
        else cudaFree(p);
    }
    void block_cache_release() {
        for (auto& b : block_cache) cudaFree(b.second);
        block_cache.clear();
        block_cache_bytes = 0;
    }
    cudaError_t arena_alloc(void** p, size_t bytes) {
        bytes = (bytes + 255) & ~(size_t)255;
        arena_peak += bytes;
        if (arena_off + bytes <= arena_cap) { *p = arena + arena_off; arena_off += bytes; return cudaSuccess; }
        cudaError_t e = cudaMalloc(p, bytes);
        if (e == cudaSuccess) arena_overflow.push_back(*p);
        return e;
    }

    int fail(int code, const std::string& msg) { last_error = msg; return code; }
    int cuda_fail(cudaError_t e, const char* where) {
        last_error = std::string(where) + ": " + cudaGetErrorString(e);
        return GPSLC_ERR_CUDA;
    }
};

// brackets one C-ABI call: resets the staging arena on entry, releases overflow allocations on exit
struct ArenaScope {
    Ctx* c;
    explicit ArenaScope(Ctx* ctx) : c(ctx) { c->arena_begin(); }
    ~ArenaScope() { c->arena_end(); }
};

#define GP_CUDA(ctx, call)                                                   \
    do {                                                                     \
        cudaError_t _e = (call);                                             \
        if (_e != cudaSuccess) return (ctx)->cuda_fail(_e, #call);           \
    } while (0)

// Reserve factor workspaces for `tasks` independent NRB x NCB block matrices and return the grid size to launch: one slot per
// resident CTA, fewer when there are fewer tasks or when the slots would not fit in free HBM (large n). With team > 1 (cluster
// launches, factor.cuh) a slot of L scratch serves a whole team and *grid is the number of teams.
int ensure_workspace(Ctx* ctx, int NRB, int NCB, long long tasks, int* grid, int team = 1, int split = 0);
// split > 0: a slot holds block rows >= split only (shared-prefix factorizations, factor.cuh PRE)
// z / w side buffers only (one per CTA) for kernels that keep their factor elsewhere; sets ctx->slot_z_d
int ensure_zbuf(Ctx* ctx, int NCB, long long ctas);
// cluster size (1, 2, 4 or 8 CTAs per matrix) for a launch with `tasks` independent factorizations of NCB block columns
int pick_team(Ctx* ctx, long long tasks, int NCB);

}  // namespace gpslc
