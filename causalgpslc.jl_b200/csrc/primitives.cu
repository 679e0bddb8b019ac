// Deterministic primitives of the C ABI (parity layer): covariance build, batched Cholesky log-density on dense
// input, and the fused build+Cholesky log-density. See include/gpslc.h for the reference call sites each replaces.
#include <cstdlib>
#include "context.cuh"
#include "gens.cuh"

namespace gpslc {

int ensure_workspace(Ctx* ctx, int NRB, int NCB, long long tasks, int* grid_out, int team, int split) {
    const size_t need = scratch_doubles(NRB, NCB) - row_off(split);
    const size_t needz = (size_t)2 * MAXRHS * NCB * NB;
    if (ctx->slots == 0) ctx->slots = CTAS_PER_SM * ctx->num_sms;   // upper bound on the grid of any factor kernel (sizes per-slot side buffers)
    // Resident factor CTAs per SM for this launch: two, so that one CTA's serial phases overlap the other's tensor work - except
    // for few large tasks (measured: 32 chains at n = 8192 run at 0.80 of peak with one CTA per SM and 0.42 with two, because the
    // second wave is nearly empty and every task runs at half speed; n = 2048, 64 chains: two CTAs 0.75, one 0.73; n = 4096, 64
    // chains: two 0.81, one 0.83). GPSLC_CTAS_PER_SM overrides (development knob).
    const char* e = getenv("GPSLC_CTAS_PER_SM");
    // Cluster teams on large matrices (measured, 32 augmented 16384 x 16384 factorizations of the n = 8192 counterfactual sweep:
    // teams of 8 with two CTAs per SM 2.3 s, teams of 4 with one CTA per SM 1.7 s): 256 CTAs on 148 SMs leave every team waiting for its
    // members that share an SM, so large-matrix teams get an SM per CTA.
    int per = (e && atoi(e) > 0) ? atoi(e)
              : ((team == 1 && ((NCB >= 32 && tasks <= 3LL * ctx->num_sms) || (NCB >= 64 && tasks <= 6LL * ctx->num_sms))) ? 1
                 : (team > 1 && NCB >= 32) ? 1 : CTAS_PER_SM);
    if (per > CTAS_PER_SM) per = CTAS_PER_SM;
    // team > 1: one L scratch per team (cluster), one z buffer per CTA; *grid_out is the number of teams
    const long long max_slots = (long long)per * ctx->num_sms / team;
    long long grid = tasks < max_slots ? tasks : max_slots;
    if (grid < 1) grid = 1;
    if ((size_t)grid * need > ctx->scratch_cap_d) {
        // bound the slot count by what fits in HBM (n = 8192 augmented: ~1 GiB per slot)
        size_t free_b = 0, total_b = 0;
        GP_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
        const size_t avail = (size_t)(0.8 * (double)(free_b + ctx->scratch_cap_d * sizeof(double)));
        long long fit = (long long)(avail / ((need + needz * team) * sizeof(double)));
        if (fit < 1) return ctx->fail(GPSLC_ERR_CUDA, "not enough device memory for one factor workspace");
        if (grid > fit) grid = fit;
    }
    if ((size_t)grid * need > ctx->scratch_cap_d) {
        if (ctx->scratch) cudaFree(ctx->scratch);
        ctx->scratch = nullptr; ctx->scratch_cap_d = 0;
        GP_CUDA(ctx, cudaMalloc(&ctx->scratch, (size_t)grid * need * sizeof(double)));
        ctx->scratch_cap_d = (size_t)grid * need;
    }
    if ((size_t)grid * team * needz > ctx->z_cap_d) {
        if (ctx->zbuf) cudaFree(ctx->zbuf);
        ctx->zbuf = nullptr; ctx->z_cap_d = 0;
        GP_CUDA(ctx, cudaMalloc(&ctx->zbuf, (size_t)grid * team * needz * sizeof(double)));
        ctx->z_cap_d = (size_t)grid * team * needz;
    }
    ctx->slot_scratch_d = need;
    ctx->slot_z_d = needz;
    if (!ctx->counter) GP_CUDA(ctx, cudaMalloc(&ctx->counter, 64 * sizeof(unsigned int)));
    *grid_out = (int)grid;
    return GPSLC_OK;
}

int ensure_zbuf(Ctx* ctx, int NCB, long long ctas) {
    const size_t needz = (size_t)2 * MAXRHS * NCB * NB;
    if ((size_t)ctas * needz > ctx->z_cap_d) {
        if (ctx->zbuf) cudaFree(ctx->zbuf);
        ctx->zbuf = nullptr; ctx->z_cap_d = 0;
        GP_CUDA(ctx, cudaMalloc(&ctx->zbuf, (size_t)ctas * needz * sizeof(double)));
        ctx->z_cap_d = (size_t)ctas * needz;
    }
    ctx->slot_z_d = needz;
    if (!ctx->counter) GP_CUDA(ctx, cudaMalloc(&ctx->counter, 64 * sizeof(unsigned int)));
    return GPSLC_OK;
}

// Team size for `tasks` independent factorizations with NCB block columns: the largest power of two (<= 8, the portable cluster
// limit) that still leaves every task a team of its own among the resident CTAs and gives every CTA at least one block row.
// The redundant diagonal work of a team is about 1.7 * team / NCB of the total. GPSLC_TEAM overrides (development knob).
int pick_team(Ctx* ctx, long long tasks, int NCB) {
    if (const char* e = getenv("GPSLC_TEAM")) { const int g = atoi(e); if (g == 1 || g == 2 || g == 4 || g == 8) return g; }
    // large matrices (>= 32 panels): one CTA per SM (see ensure_workspace), so the team size is chosen against the SM count
    const long long resident = (NCB >= 32 ? 1LL : (long long)CTAS_PER_SM) * ctx->num_sms;
    int g = 1;
    if (NCB < 8) return 1;   // below n = 512 the redundant diagonal work and the cluster barriers eat the gain
    while (g < 8 && tasks * (2 * g) <= resident && 2 * g <= NCB) g *= 2;   // SMs would idle otherwise: larger teams win even at
                                                                            // 16 panels (single chain, n = 1024: 18.4 / 18.3 / 16.6 ms
                                                                            // per sweep with 2 / 4 / 8 CTAs, 34.4 ms with one)
    return g;
}

// ------------------------------------------------------------------------------------------------ cov build
// Materialised covariance (HBM-bound): K[b] (n x n, column-major) for `batch` parameter sets. One thread computes a
// 1x4 strip of a column block so that stores are 32-byte vectors along the fastest (row) dimension.
// feat: [batch or 1][D][n] feature columns (row-contiguous), w: [batch][D] = 1 / lengthscale, scale/noise: [batch].
// X2 != X1 is supported (likelihood.jl:27 builds K(T, doT)).
// one aligned 32-byte sector per lane: 256-bit streaming store (STG.E.EF.256 on sm_100)
__device__ __forceinline__ void store4_cs(double* dst, double a, double b, double c, double d) {
    asm volatile("st.global.cs.v4.f64 [%0], {%1, %2, %3, %4};" ::"l"(dst), "d"(a), "d"(b), "d"(c), "d"(d) : "memory");
}

__global__ void __launch_bounds__(256) cov_build_kernel(int n, int D, const double* __restrict__ f1, const double* __restrict__ f2,
                                                        size_t feat_stride, const double* __restrict__ w,
                                                        const double* __restrict__ scale, const double* __restrict__ noise,
                                                        int has_noise, double* __restrict__ K) {
    // At D = 12 this kernel is bound by the FP64 pipe, not by HBM, unless the per-entry instruction count is kept low: features are
    // staged pre-scaled by 1 / lengthscale (2 instructions per dimension and entry instead of 3) and the exponential is the
    // 11-instruction table routine of gens.cuh.
    extern __shared__ double sh[];  // [32] 2^(i/32) table, then [D][32] column features, then [D][128] row features
    const int b = blockIdx.z;
    const int r0 = blockIdx.x * 128, c0 = blockIdx.y * 32;
    double* tab = sh;
    double* sc = sh + 32;           // [D][32]  features of the tile's columns (from f2), times 1 / lengthscale
    double* sr = sc + D * 32;       // [D][128] features of the tile's rows (from f1), times 1 / lengthscale
    const double* p1 = f1 + (size_t)b * feat_stride;
    const double* p2 = f2 + (size_t)b * feat_stride;
    if (threadIdx.x < 32) tab[threadIdx.x] = GPSLC_EXP2_TAB[threadIdx.x];
    for (int i = threadIdx.x; i < D * 32; i += blockDim.x) {
        const int d = i >> 5, c = c0 + (i & 31);
        sc[i] = (c < n) ? p2[(size_t)d * n + c] * w[(size_t)b * D + d] : 0.0;
    }
    for (int i = threadIdx.x; i < D * 128; i += blockDim.x) {
        const int d = i >> 7, r = r0 + (i & 127);
        sr[i] = (r < n) ? p1[(size_t)d * n + r] * w[(size_t)b * D + d] : 0.0;
    }
    __syncthreads();
    const double s = scale[b];
    const double nz = has_noise ? noise[b] : 0.0;
    // thread t: rows (t%32)*4 .. +3, columns (t/32)*4 .. +3
    const int tr = (threadIdx.x & 31) * 4, tc = (threadIdx.x >> 5) * 4;
    double acc[4][4];
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) acc[j][i] = 0.0;
    for (int d = 0; d < D; d++) {
        double zr[4], zc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) zr[i] = sr[d * 128 + tr + i];
#pragma unroll
        for (int j = 0; j < 4; j++) zc[j] = sc[d * 32 + tc + j];
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double t = zr[i] - zc[j];
                acc[j][i] = fma(t, t, acc[j][i]);
            }
    }
    double* Kb = K + (size_t)b * n * n;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c = c0 + tc + j;
        if (c >= n) continue;
        double v[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            v[i] = fma(s, exp_neg_tab(acc[j][i], tab), (has_noise && (r0 + tr + i) == c) ? nz : 0.0);
        }
        const int r = r0 + tr;
        double* dst = Kb + (size_t)c * n + r;
        if (r + 3 < n && ((((size_t)c * n + r) & 3) == 0) && ((reinterpret_cast<uintptr_t>(Kb) & 31) == 0)) {
            store4_cs(dst, v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (r + i < n) dst[i] = v[i];
        }
    }
}

// K(X, X): the matrix is symmetric, so only the 64 x 64 tiles on and below the block diagonal are computed (half the FP64 work:
// at D ~ 12 an entry costs 2 D + 12 FP64-pipe instructions, and the FP64 pipe is the second roof of this kernel) and every
// off-diagonal tile is written twice, as it is and transposed, straight from the 4 x 4 register block of a thread with 256-bit
// streaming stores (STG.E.EF.256, sm_100): every store instruction writes one full, aligned 32-byte sector per lane, so L2 never sees
// a partial sector - ncu on the first version of this kernel, which wrote each sector as two 128-bit halves, showed the L1/LSU pipe
// at 94 % and L2 at 66 % with DRAM at 36-50 % (profiles/ncu_cov_build_r02.md). The direct store is contiguous along a column
// (16 lanes x 32 B), the transposed one scatters full sectors that L2 assembles into lines. No shared-memory transpose, no barrier
// after the feature staging, 12.5 KB of shared memory and 64 registers: four CTAs per SM. Feature rows are staged permuted
// (row r at (r & 3) * 16 + (r >> 2)) so that the four row values of a thread come from four conflict-free LDS.64.
// K comes out exactly symmetric.
__global__ void __launch_bounds__(256, 4) cov_build_sym_kernel(int n, int D, const double* __restrict__ f, size_t feat_stride,
                                                               const double* __restrict__ w, const double* __restrict__ scale,
                                                               const double* __restrict__ noise, int has_noise, double* __restrict__ K) {
    extern __shared__ __align__(16) double sh[];  // [32] 2^(i/32) table, [D][64] column features, [D][64] row features (permuted)
    const int b = blockIdx.z;
    // linear index of a lower-triangular tile pair -> (bi >= bj)
    const int p = blockIdx.x;
    int bi = (int)((sqrt(8.0 * p + 1.0) - 1.0) * 0.5);
    while ((bi + 1) * (bi + 2) / 2 <= p) bi++;
    while (bi * (bi + 1) / 2 > p) bi--;
    const int bj = p - bi * (bi + 1) / 2;
    const int r0 = bi * 64, c0 = bj * 64;
    double* tab = sh;
    double* sc = sh + 32;
    double* sr = sc + D * 64;
    const double* pf = f + (size_t)b * feat_stride;
    if (threadIdx.x < 32) tab[threadIdx.x] = GPSLC_EXP2_TAB[threadIdx.x];
    for (int i = threadIdx.x; i < D * 64; i += blockDim.x) {
        const int d = i >> 6, o = i & 63;
        const int po = d * 64 + (o & 3) * 16 + (o >> 2);      // permuted position inside the dimension's 64 values
        const double sw = w[(size_t)b * D + d];     // 1 / lengthscale
        sc[po] = (c0 + o < n) ? pf[(size_t)d * n + c0 + o] * sw : 0.0;
        sr[po] = (r0 + o < n) ? pf[(size_t)d * n + r0 + o] * sw : 0.0;
    }
    __syncthreads();
    const double s = scale[b];
    const double nz = has_noise ? noise[b] : 0.0;
    const int lr = threadIdx.x & 15, lc = threadIdx.x >> 4;
    const int tr = lr * 4, tc = lc * 4;
    double acc[4][4];       // [column j][row i]
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++) acc[j][i] = 0.0;
#pragma unroll 2
    for (int d = 0; d < D; d++) {
        double zr[4], zc[4];
#pragma unroll
        for (int i = 0; i < 4; i++) { zr[i] = sr[d * 64 + i * 16 + lr]; zc[i] = sc[d * 64 + i * 16 + lc]; }
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const double t = zr[i] - zc[j];
                acc[j][i] = fma(t, t, acc[j][i]);
            }
    }
#pragma unroll
    for (int j = 0; j < 4; j++)
#pragma unroll
        for (int i = 0; i < 4; i++)
            acc[j][i] = fma(s, exp_neg_tab(acc[j][i], tab), (has_noise && (r0 + tr + i) == (c0 + tc + j)) ? nz : 0.0);
    double* Kb = K + (size_t)b * n * n;
    const bool vec_ok = ((n & 3) == 0) && ((reinterpret_cast<uintptr_t>(Kb) & 31) == 0);
    // direct: column c0+tc+j, rows r0+tr .. +3
#pragma unroll
    for (int j = 0; j < 4; j++) {
        const int c = c0 + tc + j, r = r0 + tr;
        if (c >= n) continue;
        double* dst = Kb + (size_t)c * n + r;
        if (vec_ok && r + 3 < n) store4_cs(dst, acc[j][0], acc[j][1], acc[j][2], acc[j][3]);
        else {
#pragma unroll
            for (int i = 0; i < 4; i++)
                if (r + i < n) dst[i] = acc[j][i];
        }
    }
    if (bi == bj) return;
    // transposed: column r0+tr+i (a row of this tile), rows c0+tc .. +3 (its columns)
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const int cc = r0 + tr + i, rr = c0 + tc;
        if (cc >= n) continue;
        double* dst = Kb + (size_t)cc * n + rr;
        if (vec_ok && rr + 3 < n) store4_cs(dst, acc[0][i], acc[1][i], acc[2][i], acc[3][i]);
        else {
#pragma unroll
            for (int jj = 0; jj < 4; jj++)
                if (rr + jj < n) dst[jj] = acc[jj][i];
        }
    }
}

int launch_cov_build(Ctx* ctx, int n, int batch, int D, const double* f1, const double* f2, size_t feat_stride,
                     const double* w, const double* scale, const double* noise, double* K) {
    if (f1 == f2) {
        const int nt = ceil_div(n, 64);
        dim3 grid(nt * (nt + 1) / 2, 1, batch);
        const size_t sh = (size_t)(32 + 2 * D * 64) * sizeof(double);
        if (sh > 48 * 1024) GP_CUDA(ctx, cudaFuncSetAttribute(cov_build_sym_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        cov_build_sym_kernel<<<grid, 256, sh, ctx->stream>>>(n, D, f1, feat_stride, w, scale, noise, noise != nullptr, K);
        ctx->launches++;
        GP_CUDA(ctx, cudaGetLastError());
        return GPSLC_OK;
    }
    dim3 grid(ceil_div(n, 128), ceil_div(n, 32), batch);
    size_t sh = (size_t)(32 + D * 32 + D * 128) * sizeof(double);
    if (sh > 48 * 1024) GP_CUDA(ctx, cudaFuncSetAttribute(cov_build_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    cov_build_kernel<<<grid, 256, sh, ctx->stream>>>(n, D, f1, f2, feat_stride, w, scale, noise, noise != nullptr, K);
    ctx->launches++;
    GP_CUDA(ctx, cudaGetLastError());
    return GPSLC_OK;
}

// ------------------------------------------------------------------------------------------------ batched logpdf
struct LogpdfJobDense { const double* K; const double* y; int n; int ld; };

template <bool FUSED>
__global__ void __launch_bounds__(FTHREADS, CTAS_PER_SM)
batched_logpdf_kernel(int batch, int n, const double* __restrict__ Kall, int ld, const double* __restrict__ yall, int y_shared,
                      // fused build inputs
                      int D, const double* __restrict__ feat, size_t feat_stride, const double* __restrict__ w,
                      const double* __restrict__ scale, const double* __restrict__ noise,
                      double* scratch, size_t slot_scratch, double* zbuf, size_t slot_z, unsigned int* counter,
                      double* logpdf, double* logdet_out, double* quad_out, int* info) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    __shared__ RbfSpec spec;
    __shared__ unsigned int job;
    factor_smem_init(sm);
    Pipe pipe{0, 0};
    const int NCB = ceil_div(n, NB);
    double* my_scratch = scratch + (size_t)blockIdx.x * slot_scratch;
    double* my_z = zbuf + (size_t)blockIdx.x * slot_z;
    for (;;) {
        if (threadIdx.x == 0) job = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned int b = job;
        if (b >= (unsigned)batch) break;
        const double* y = yall + (y_shared ? 0 : (size_t)b * n);
        if (FUSED) {
            if (threadIdx.x == 0) {
                spec.D = D; spec.n = n; spec.scale = scale[b]; spec.noise = noise[b];
                spec.y[0] = y; spec.y[1] = y;
            }
            for (int d = threadIdx.x; d < D; d += blockDim.x) {
                spec.feat[d] = feat + (size_t)b * feat_stride + (size_t)d * n;
                spec.w[d] = w[(size_t)b * D + d];
                spec.sw[d] = sqrt(spec.w[d]);
            }
            __syncthreads();
            RbfGen gen{&spec, sm.exp2tab};
            factor_run(gen, NCB, NCB, 1, my_scratch, my_z, sm, pipe);
        } else {
            DenseGen gen{Kall + (size_t)b * ld * n, {y, y}, n, ld};
            factor_run(gen, NCB, NCB, 1, my_scratch, my_z, sm, pipe);
        }
        if (threadIdx.x == 0) {
            const FactorOut o = sm.out;
            if (logdet_out) logdet_out[b] = o.logdet;
            if (quad_out) quad_out[b] = o.gram[0];
            if (info) info[b] = o.info;
            if (logpdf) logpdf[b] = (o.info == 0) ? -0.5 * (n * LOG_2PI + o.logdet + o.gram[0]) : -INFINITY;
        }
        __syncthreads();
    }
}

template <bool FUSED>
static int launch_batched(Ctx* ctx, int batch, int n, const double* K, int ld, const double* y, int y_shared, int D,
                          const double* feat, size_t feat_stride, const double* w, const double* scale,
                          const double* noise, double* logpdf, double* logdet, double* quad, int* info) {
    const int NCB = ceil_div(n, NB);
    int grid = 0;
    int rc = ensure_workspace(ctx, NCB, NCB, batch, &grid);
    if (rc) return rc;
    auto kern = batched_logpdf_kernel<FUSED>;
    GP_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
    GP_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
    kern<<<grid, FTHREADS, sizeof(FactorSmem), ctx->stream>>>(batch, n, K, ld, y, y_shared, D, feat, feat_stride, w, scale,
                                                             noise, ctx->scratch, ctx->slot_scratch_d, ctx->zbuf,
                                                             ctx->slot_z_d, ctx->counter, logpdf, logdet, quad, info);
    ctx->launches++;
    GP_CUDA(ctx, cudaGetLastError());
    return GPSLC_OK;
}

int launch_chol_logpdf_dense(Ctx* ctx, int batch, int n, const double* K, int ld, const double* y, int y_shared,
                             double* logpdf, double* logdet, double* quad, int* info) {
    return launch_batched<false>(ctx, batch, n, K, ld, y, y_shared, 0, nullptr, 0, nullptr, nullptr, nullptr, logpdf, logdet,
                                 quad, info);
}
int launch_rbf_logpdf(Ctx* ctx, int batch, int n, int D, const double* feat, size_t feat_stride, const double* w,
                      const double* scale, const double* noise, const double* y, int y_shared, double* logpdf,
                      double* logdet, double* quad, int* info) {
    return launch_batched<true>(ctx, batch, n, nullptr, 0, y, y_shared, D, feat, feat_stride, w, scale, noise, logpdf, logdet,
                                quad, info);
}

}  // namespace gpslc
