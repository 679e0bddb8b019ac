// Posterior-predictive treatment effects: the GP conditional of src/likelihood.jl:8-174 and src/estimation.jl:36-163
// restructured as ONE Cholesky per (posterior sample, doT):
//
//   A = [ Kp        D            ]   Kp = Kww + yNoise I,   D = Kws - Kww,
//       [ D'   P + jitter I      ]   P  = Kww - Kws - Kws' + Kss
//
//   chol(A) = [ L11 0 ; W' L22 ],  W = L11^-1 D,  L22 L22' = P - W'W + jitter I = Symmetric(CovITE) + jitter I
//   forward solve with rhs [Y; 0]: z1 = L11^-1 Y, and the pre-solve residual of the second block is -W'z1 = -MeanITE.
//
// The reference instead factorises Kp three times (LU, LU, Bunch-Kaufman), forms four n x n products and re-factorises
// CovITE for every draw (src/likelihood.jl:42-49, src/estimation.jl:46,105).
// SATE only needs 1'MeanITE and 1'CovITE 1 (src/estimation.jl:116-121): one n x n Cholesky with right-hand sides
// (Y, D1) — gpslc_sate never forms CovITE.
#include "context.cuh"
#include "gens.cuh"
#include "rng.cuh"
#include "capi_util.cuh"
#include "est.cuh"

namespace gpslc {

struct IteSpec {
    const double* feat[DMAX];   // U columns (record, column-wise as extractParameters builds them) then X columns
    double w[DMAX];
    const double* T;
    const double* Y;
    double wT, doT, yScale, yNoise, jitter;
    int D, n, npad;
};

struct IteGen {
    const IteSpec* s;
    __device__ __forceinline__ double base(int i, int j) const {
        double a = 0.0;
        for (int d = 0; d < s->D; d++) {
            const double* p = s->feat[d];
            const double t = p[i] - p[j];
            a = fma(t * s->w[d], t, a);
        }
        return a;
    }
    __device__ __forceinline__ double one(int r, int c) const {
        const int n = s->n, np = s->npad;
        const bool r2 = r >= np, c2 = c >= np;
        const int i = r2 ? r - np : r, j = c2 ? c - np : c;
        if (i >= n || j >= n) return (r == c) ? 1.0 : 0.0;
        const double b = base(i, j);
        const double ti = s->T[i], tj = s->T[j];
        const double dt = ti - tj, tt = dt * s->wT * dt;
        if (!r2) {  // Kp
            double v = s->yScale * exp(-(b + tt));
            if (i == j) v += s->yNoise;
            return v;
        }
        const double di = ti - s->doT, dj = tj - s->doT;
        if (!c2) {  // D'[i][j] = Kws[j][i] - Kww[j][i]
            return s->yScale * exp(-(b + dj * s->wT * dj)) - s->yScale * exp(-(b + tt));
        }
        // P + jitter I
        double v = s->yScale * exp(-(b + tt)) - s->yScale * exp(-(b + di * s->wT * di)) - s->yScale * exp(-(b + dj * s->wT * dj)) +
                   s->yScale * exp(-b);
        if (i == j) v += s->jitter;
        return v;
    }
    __device__ __forceinline__ void quad(int r0, int r1, int c, double& v00, double& v01, double& v10, double& v11) const {
        v00 = one(r0, c); v01 = one(r0, c + 1); v10 = one(r1, c); v11 = one(r1, c + 1);
    }
    __device__ __forceinline__ double rhs(int which, int r) const { return (r < s->n) ? s->Y[r] : 0.0; }
    GPSLC_GENERIC_STRIP
};


__device__ inline void fill_ite_spec(const EstArgs& a, const double* rec, double doT, IteSpec* sp) {
    // record layout
    const int nX = a.nX, nU = a.nU;
    for (int d = threadIdx.x; d < nU + nX; d += blockDim.x) {
        if (d < nU) {
            sp->feat[d] = rec + a.n_params + (size_t)d * a.n;
            const double ls = rec[6 + 4 * nX + nU + d];           // uyLS
            sp->w[d] = 1.0 / (ls * ls);
        } else {
            const int k = d - nU;
            sp->feat[d] = a.X + (size_t)k * a.n;
            const double ls = rec[6 + 3 * nX + k];                // xyLS
            sp->w[d] = 1.0 / (ls * ls);
        }
    }
    if (threadIdx.x == 0) {
        sp->D = nU + nX; sp->n = a.n; sp->npad = ceil_div(a.n, NB) * NB;
        sp->T = a.T; sp->Y = a.Y;
        const double tyLS = rec[3];
        sp->wT = 1.0 / (tyLS * tyLS);
        sp->doT = doT; sp->yNoise = rec[2]; sp->yScale = rec[5]; sp->jitter = a.jitter;
    }
}

// one task = (doT d, chain c, retained sample r): augmented Cholesky, MeanITE, optional CovITE, ITE draws
__global__ void __launch_bounds__(FTHREADS, 2)
ite_kernel(EstArgs a, double* scratch, size_t slot_scratch, double* zbuf, size_t slot_z, double* xibuf, unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    __shared__ IteSpec spec;
    __shared__ unsigned int job;
    factor_smem_init(sm);
    Pipe pipe{0, 0};
    const int NCB1 = ceil_div(a.n, NB), NCB = 2 * NCB1, npad = NCB1 * NB;
    double* my_scratch = scratch + (size_t)blockIdx.x * slot_scratch;
    double* my_z = zbuf + (size_t)blockIdx.x * slot_z;
    double* xi = xibuf + (size_t)blockIdx.x * 4 * a.n;
    const unsigned int total = (unsigned)a.n_doT * a.n_chains * a.R;
    for (;;) {
        if (threadIdx.x == 0) job = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned int t = job;
        if (t >= total) break;
        const int r = t % a.R, c = (t / a.R) % a.n_chains, d = t / (a.R * a.n_chains);
        const double* rec = a.samples + ((size_t)a.ret_idx[r] * a.n_chains + c) * a.stride;
        fill_ite_spec(a, rec, a.doT[d], &spec);
        __syncthreads();
        IteGen gen{&spec};
        double* cov = a.cov_out ? a.cov_out + (size_t)t * a.n * a.n : nullptr;
        factor_run(gen, NCB, NCB, 1, my_scratch, my_z, sm, pipe, NCB1, cov, a.n);
        const int info = sm.out.info;
        if (threadIdx.x == 0 && a.info) a.info[t] = info;
        // MeanITE = -(pre-solve residual of the second block)
        const double* wres = my_z + (size_t)MAXRHS * (2 * npad) + npad;
        if (a.mean_out)
            for (int i = threadIdx.x; i < a.n; i += blockDim.x) a.mean_out[(size_t)t * a.n + i] = -wres[i];
        // draws: MeanITE + L22 xi   (src/estimation.jl:95-109; one factor serves all spp draws)
        if (a.ite_out && a.spp > 0) {
            const unsigned gchain = (unsigned)(a.chain0 + c);
            for (int s0 = 0; s0 < a.spp; s0 += 4) {
                const int ns = min(4, a.spp - s0);
                __syncthreads();
                for (int e = threadIdx.x; e < ns * a.n; e += blockDim.x) {
                    const int s = e / a.n, col = e - s * a.n;
                    Stream st(a.seed, gchain, (uint32_t)(r * a.spp + s0 + s), stream_b(TAG_ITE, (uint32_t)d));
                    xi[(size_t)s * a.n + col] = st.normal_at(col);
                }
                __syncthreads();
                for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
                    double accv[4] = {0.0, 0.0, 0.0, 0.0};
                    const int I = NCB1 + (i >> 6), ri = i & 63;
                    for (int jb = 0; jb <= (i >> 6); jb++) {
                        const double* blk = my_scratch + block_off(I, NCB1 + jb, NCB);
                        const int jmax = (jb == (i >> 6)) ? ri : 63;
                        for (int jj = 0; jj <= jmax; jj++) {
                            const double l = blk[elem_off(ri, jj)];
                            const int col = jb * 64 + jj;
                            if (col < a.n) {
#pragma unroll
                                for (int s = 0; s < 4; s++)
                                    if (s < ns) accv[s] = fma(l, xi[(size_t)s * a.n + col], accv[s]);
                            }
                        }
                    }
                    for (int s = 0; s < ns; s++)
                        a.ite_out[(((size_t)d * a.n_chains + c) * a.R * a.spp + (size_t)r * a.spp + s0 + s) * a.n + i] = -wres[i] + accv[s];
                }
            }
        }
        __syncthreads();
    }
}

// SATE fast path: one task = (doT, chain, retained sample)
__global__ void __launch_bounds__(FTHREADS, 2)
sate_kernel(EstArgs a, double* scratch, size_t slot_scratch, double* zbuf, size_t slot_z, double* d1buf, unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    __shared__ IteSpec spec;
    __shared__ RbfSpec rspec;
    __shared__ unsigned int job;
    __shared__ double red[32];
    factor_smem_init(sm);
    Pipe pipe{0, 0};
    const int NCB = ceil_div(a.n, NB);
    double* my_scratch = scratch + (size_t)blockIdx.x * slot_scratch;
    double* my_z = zbuf + (size_t)blockIdx.x * slot_z;
    double* d1 = d1buf + (size_t)blockIdx.x * a.n;
    const unsigned int total = (unsigned)a.n_doT * a.n_chains * a.R;
    for (;;) {
        if (threadIdx.x == 0) job = atomicAdd(counter, 1u);
        __syncthreads();
        const unsigned int t = job;
        if (t >= total) break;
        const int r = t % a.R, c = (t / a.R) % a.n_chains, d = t / (a.R * a.n_chains);
        const double* rec = a.samples + ((size_t)a.ret_idx[r] * a.n_chains + c) * a.stride;
        fill_ite_spec(a, rec, a.doT[d], &spec);
        __syncthreads();
        // Row sums of Kww and Kss give D1 = (Kws - Kww) 1 = a .* rs_ss - rs_ww  and
        // 1'P1 = sum(rs_ww) - 2 a.rs_ss + sum(rs_ss), with a_i = exp(-(T_i - doT)^2 / tyLS^2)  
        // One row per thread, four columns per step with independent accumulators and the branch-free exp, so the FP64 pipe
        // is throughput- rather than latency-bound; column features are warp-uniform (broadcast) loads.
        double psum = 0.0;
        for (int i = threadIdx.x; i < a.n; i += blockDim.x) {
            double ww[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
            const double ti = a.T[i];
            const int D = spec.D;
            for (int j0 = 0; j0 < a.n; j0 += 4) {
                double b[4] = {0.0, 0.0, 0.0, 0.0};
                for (int d = 0; d < D; d++) {
                    const double* p = spec.feat[d];
                    const double w = spec.w[d];
                    const double zi = __ldg(p + i);
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int j = min(j0 + u, a.n - 1);
                        const double t = zi - __ldg(p + j);
                        b[u] = fma(t * w, t, b[u]);
                    }
                }
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    if (j0 + u < a.n) {
                        const double dt = ti - __ldg(a.T + j0 + u);
                        ww[u] += spec.yScale * exp_neg(b[u] + dt * spec.wT * dt);
                        ss[u] += spec.yScale * exp_neg(b[u]);
                    }
                }
            }
            const double rs_ww = (ww[0] + ww[1]) + (ww[2] + ww[3]), rs_ss = (ss[0] + ss[1]) + (ss[2] + ss[3]);
            const double di = ti - spec.doT;
            const double ai = exp(-(di * spec.wT * di));
            d1[i] = ai * rs_ss - rs_ww;
            psum += rs_ww - 2.0 * ai * rs_ss + rs_ss;
        }
        {
            double v = psum;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            double v = 0.0;
            for (int w = 0; w < FWARPS; w++) v += red[w];
            red[16] = v;
            // Kp as an RBF factor with the T dimension appended
            rspec.D = spec.D + 1; rspec.n = a.n; rspec.scale = spec.yScale; rspec.noise = spec.yNoise;
            rspec.y[0] = a.Y; rspec.y[1] = d1;
            rspec.feat[spec.D] = a.T; rspec.w[spec.D] = spec.wT;
        }
        for (int k = threadIdx.x; k < spec.D; k += blockDim.x) { rspec.feat[k] = spec.feat[k]; rspec.w[k] = spec.w[k]; }
        __syncthreads();
        RbfGen rgen{&rspec};
        factor_run(rgen, NCB, NCB, 2, my_scratch, my_z, sm, pipe);
        if (threadIdx.x == 0) {
            const FactorOut o = sm.out;
            const double n = (double)a.n;
            const double ms = o.gram[1] / n;                                   // 1'D'Kp^-1 Y / n
            const double vs = (red[16] - o.gram[2]) / (n * n) + a.jitter / n;  // 1'(CovITE + jitter I)1 / n^2
            if (a.info) a.info[t] = o.info;
            if (a.msate) a.msate[t] = ms;
            if (a.vsate) a.vsate[t] = vs;
            if (a.sate_out) {
                const unsigned gchain = (unsigned)(a.chain0 + c);
                // normal(mean, var): Gen's second argument is a std, the reference passes the variance (App. B5)
                const double sd = a.var_as_std ? vs : sqrt(fmax(vs, 0.0));
                for (int s = 0; s < a.spp; s++) {
                    Stream st(a.seed, gchain, (uint32_t)(r * a.spp + s), stream_b(TAG_SATE, (uint32_t)d));
                    a.sate_out[((size_t)d * a.n_chains + c) * a.R * a.spp + (size_t)r * a.spp + s] = ms + sd * st.normal();
                }
            }
        }
        __syncthreads();
    }
}

int launch_ite(Ctx* ctx, const EstArgs& a) {
    const int NCB = 2 * ceil_div(a.n, NB);
    int grid = 0;
    GP_TRY(ensure_workspace(ctx, NCB, NCB, (long long)a.n_doT * a.n_chains * a.R, &grid));
    GP_CUDA(ctx, cudaFuncSetAttribute(ite_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
    GP_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
    const long long total = (long long)a.n_doT * a.n_chains * a.R;
    if (total == 0) return GPSLC_OK;
    double* xi = nullptr;
    GP_CUDA(ctx, cudaMalloc(&xi, (size_t)grid * 4 * a.n * sizeof(double)));
    ite_kernel<<<grid, FTHREADS, sizeof(FactorSmem), ctx->stream>>>(a, ctx->scratch, ctx->slot_scratch_d, ctx->zbuf, ctx->slot_z_d, xi, ctx->counter);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(xi);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "ite_kernel");
    if (e2 != cudaSuccess) return ctx->cuda_fail(e2, "ite_kernel");
    return GPSLC_OK;
}

int launch_sate(Ctx* ctx, const EstArgs& a) {
    const int NCB = ceil_div(a.n, NB);
    int grid = 0;
    GP_TRY(ensure_workspace(ctx, NCB, NCB, (long long)a.n_doT * a.n_chains * a.R, &grid));
    GP_CUDA(ctx, cudaFuncSetAttribute(sate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
    GP_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
    const long long total = (long long)a.n_doT * a.n_chains * a.R;
    if (total == 0) return GPSLC_OK;
    double* d1 = nullptr;
    GP_CUDA(ctx, cudaMalloc(&d1, (size_t)grid * a.n * sizeof(double)));
    sate_kernel<<<grid, FTHREADS, sizeof(FactorSmem), ctx->stream>>>(a, ctx->scratch, ctx->slot_scratch_d, ctx->zbuf, ctx->slot_z_d, d1, ctx->counter);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(d1);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "sate_kernel");
    if (e2 != cudaSuccess) return ctx->cuda_fail(e2, "sate_kernel");
    return GPSLC_OK;
}

}  // namespace gpslc
