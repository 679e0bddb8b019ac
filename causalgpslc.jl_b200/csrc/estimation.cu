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
    double sw[DMAX];            // 1 / lengthscale: the tile generator works on pre-scaled features (see RbfGen)
    const double* T;
    const double* Y;
    double wT, swT, doT, yScale, yNoise, jitter;
    int D, n, npad;
};

// The strip() fast path and the element-wise one() (ragged edges, snapshot of diagonal tiles) must give bit-identical values for
// the same entry: which of the two evaluates an entry depends on how the row blocks are tiled (one CTA vs cluster teams), and the
// results are required not to. Both therefore go through the helpers below, written with explicit round-to-nearest intrinsics so
// that the compiler cannot contract a multiply-add in one path and not in the other.
struct IteGen {
    const IteSpec* s;
    const double* tab;          // FactorSmem::exp2tab
    // exp(-(t - doT)^2 / tyLS^2)
    __device__ __forceinline__ double a_of(double t) const {
        const double d = __dsub_rn(t, s->doT);
        return exp_neg_tab(__dmul_rn(__dmul_rn(d, s->wT), d), tab);
    }
    // entry value from the shared-dimension distance b, the scaled treatments tis/tjs, and a_i, a_j
    __device__ __forceinline__ double entry(double b, double tis, double tjs, double ai, double aj, bool r2, bool c2, bool on_diag) const {
        const double dt = __dsub_rn(tis, tjs);
        const double eij = exp_neg_tab(__dmul_rn(dt, dt), tab);
        double gfac;
        if (!r2) gfac = eij;                                                     // Kp
        else if (!c2) gfac = __dsub_rn(aj, eij);                                 // D'[i][j] = Kws[j][i] - Kww[j][i]
        else gfac = __dadd_rn(__dsub_rn(__dsub_rn(eij, ai), aj), 1.0);           // P
        double val = __dmul_rn(__dmul_rn(s->yScale, exp_neg_tab(b, tab)), gfac);
        if (on_diag) val = __dadd_rn(val, c2 ? s->jitter : s->yNoise);
        return val;
    }
    __device__ __noinline__ double one(int r, int c) const {      // ragged edges and the diagonal snapshot only (code size, see RbfGen::one)
        const int n = s->n, np = s->npad;
        const bool r2 = r >= np, c2 = c >= np;
        const int i = r2 ? r - np : r, j = c2 ? c - np : c;
        if (i >= n || j >= n) return (r == c) ? 1.0 : 0.0;
        double b = 0.0;
        for (int d = 0; d < s->D; d++) {
            const double* p = s->feat[d];
            const double t = __dsub_rn(__dmul_rn(p[i], s->sw[d]), __dmul_rn(p[j], s->sw[d]));
            b = __fma_rn(t, t, b);
        }
        const double ti = s->T[i], tj = s->T[j];
        const double ai = (r2 && c2) ? a_of(ti) : 0.0, aj = r2 ? a_of(tj) : 0.0;
        return entry(b, __dmul_rn(ti, s->swT), __dmul_rn(tj, s->swT), ai, aj, r2, c2, r == c);
    }
    __device__ __forceinline__ void quad(int r0, int r1, int c, double& v00, double& v01, double& v10, double& v11) const {
        v00 = one(r0, c); v01 = one(r0, c + 1); v10 = one(r1, c); v11 = one(r1, c + 1);
    }
    __device__ __forceinline__ double rhs(int which, int r) const { return (r < s->n) ? s->Y[r] : 0.0; }
    // Column features of the panel (zero beyond n): the D shared dimensions and T_j, all pre-scaled by 1 / lengthscale, then
    // a_j = exp(-(T_j - doT)^2 / tyLS^2).
    __device__ __forceinline__ void stage_cols(int col0, double* cf) const {
        const int D = s->D;
        if (D + 2 > CF_DIMS) return;
        const int j0 = (col0 >= s->npad) ? col0 - s->npad : col0;
        for (int i = threadIdx.x; i < (D + 2) * NB; i += blockDim.x) {
            const int d = i >> 6, c = j0 + (i & 63);
            double v = 0.0;
            if (c < s->n) {
                if (d < D) v = __dmul_rn(__ldg(s->feat[d] + c), s->sw[d]);
                else {
                    v = __ldg(s->T + c);
                    v = (d == D + 1) ? a_of(v) : __dmul_rn(v, s->swT);
                }
            }
            cf[i] = v;
        }
    }
    // Every entry of the augmented matrix is yScale * exp(-b_ij) * g, with b the shared-dimension distance and
    //   Kp: g = e_ij            D': g = a_j - e_ij            P: g = e_ij - a_i - a_j + 1,     e_ij = exp(-(T_i - T_j)^2 / tyLS^2)
    // (src/likelihood.jl:24-28 builds Kww, Kws, Kss from the same five log-kernels), so a strip costs two exponentials per entry.
    // A strip never straddles the block boundary (npad is a multiple of the panel width).
    template <int NI, bool ONE_ROW>
    __device__ __forceinline__ void strip(int r0, int r1, int c0, double (&v)[2][NI][2], const double* cf, int cl) const {
        const int n = s->n, np = s->npad, D = s->D;
        const bool r2 = r0 >= np, c2 = c0 >= np;
        const int i0 = r2 ? r0 - np : r0, i1 = r2 ? r1 - np : r1, j0 = c2 ? c0 - np : c0;
        if (D + 2 <= CF_DIMS && i0 < n && (ONE_ROW || i1 < n) && j0 + 8 * (NI - 1) + 1 < n) {
            double a[2][NI][2];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) { a[0][ni][0] = 0.0; a[0][ni][1] = 0.0; a[1][ni][0] = 0.0; a[1][ni][1] = 0.0; }
#pragma unroll 4
            for (int d = 0; d < D; d++) {
                const double* p = s->feat[d];
                const double w = s->sw[d];
                const double z0 = __dmul_rn(__ldg(p + i0), w);
                const double z1 = ONE_ROW ? z0 : __dmul_rn(__ldg(p + i1), w);
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
                    const double2 cc = *reinterpret_cast<const double2*>(cf + d * NB + cl + 8 * ni);
                    double t;
                    t = __dsub_rn(z0, cc.x); a[0][ni][0] = __fma_rn(t, t, a[0][ni][0]);
                    t = __dsub_rn(z0, cc.y); a[0][ni][1] = __fma_rn(t, t, a[0][ni][1]);
                    if (!ONE_ROW) {
                        t = __dsub_rn(z1, cc.x); a[1][ni][0] = __fma_rn(t, t, a[1][ni][0]);
                        t = __dsub_rn(z1, cc.y); a[1][ni][1] = __fma_rn(t, t, a[1][ni][1]);
                    }
                }
            }
            const double t0 = __ldg(s->T + i0), t1 = ONE_ROW ? t0 : __ldg(s->T + i1);
            const double t0s = __dmul_rn(t0, s->swT), t1s = __dmul_rn(t1, s->swT);
            double ai0 = 0.0, ai1 = 0.0;
            if (r2 && c2) { ai0 = a_of(t0); ai1 = ONE_ROW ? ai0 : a_of(t1); }
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                const double2 tc = *reinterpret_cast<const double2*>(cf + D * NB + cl + 8 * ni);
                const double2 ac = *reinterpret_cast<const double2*>(cf + (D + 1) * NB + cl + 8 * ni);
#pragma unroll
                for (int rr = 0; rr < (ONE_ROW ? 1 : 2); rr++) {
                    const int r = rr ? r1 : r0;
#pragma unroll
                    for (int e = 0; e < 2; e++)
                        v[rr][ni][e] = entry(a[rr][ni][e], rr ? t1s : t0s, e ? tc.y : tc.x, rr ? ai1 : ai0, e ? ac.y : ac.x, r2, c2,
                                             r == c0 + 8 * ni + e);
                }
            }
        } else {
#pragma unroll
            for (int ni = 0; ni < NI; ni++)
                quad(r0, r1, c0 + 8 * ni, v[0][ni][0], v[0][ni][1], v[1][ni][0], v[1][ni][1]);
        }
    }
};


__device__ inline void fill_ite_spec(const EstArgs& a, const double* rec, double doT, IteSpec* sp) {
    // record layout
    const int nX = a.nX, nU = a.nU;
    for (int d = threadIdx.x; d < nU + nX; d += blockDim.x) {
        if (d < nU) {
            sp->feat[d] = rec + a.n_params + (size_t)d * a.n;
            const double ls = rec[6 + 4 * nX + nU + d];           // uyLS
            sp->w[d] = a.ls_unsquared ? 1.0 / ls : 1.0 / (ls * ls); sp->sw[d] = a.ls_unsquared ? rsqrt(ls) : 1.0 / ls;
        } else {
            const int k = d - nU;
            sp->feat[d] = a.X + (size_t)k * a.n;
            const double ls = rec[6 + 3 * nX + k];                // xyLS
            sp->w[d] = a.ls_unsquared ? 1.0 / ls : 1.0 / (ls * ls); sp->sw[d] = a.ls_unsquared ? rsqrt(ls) : 1.0 / ls;
        }
    }
    if (threadIdx.x == 0) {
        sp->D = nU + nX; sp->n = a.n; sp->npad = ceil_div(a.n, NB) * NB;
        sp->T = a.T; sp->Y = a.Y;
        const double tyLS = rec[3];
        sp->wT = a.ls_unsquared ? 1.0 / tyLS : 1.0 / (tyLS * tyLS); sp->swT = a.ls_unsquared ? rsqrt(tyLS) : 1.0 / tyLS;
        sp->doT = doT; sp->yNoise = rec[2]; sp->yScale = rec[5]; sp->jitter = a.jitter;
    }
}

// one task = (doT d, chain c, retained sample r): augmented Cholesky, MeanITE, optional CovITE, ITE draws.
// TEAM = 0: one CTA per task, tasks handed out from an atomic counter. TEAM = 1: one thread-block cluster per task (few large
// tasks, c5 of BASELINE.json); scratch is indexed by cluster, z / xi buffers by CTA, tasks are dealt out round-robin.
// chol(Kp), L11^-1 Y and the P2 outputs of every panel, once per posterior sample (base task b = chain * R + r): what ite_kernel<.., true>
// continues from for every doT value. TEAM = 1: one cluster per base task; TEAM = 2: the whole cooperatively launched grid on one base
// task after the other (few large samples: the n = 8192 sweep has ONE).
template <int TEAM>
__global__ void __launch_bounds__(FTHREADS, CTAS_PER_SM)
ite_base_kernel(EstArgs a, double* base_L, double* base_linv, double* base_z, int* base_info, double* zbuf, size_t slot_z,
                unsigned int* counter, unsigned int* gbar) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    __shared__ IteSpec spec;
    __shared__ unsigned int job;
    factor_smem_init(sm);
    if (threadIdx.x == 0) sm.gbar = gbar;
    Pipe pipe{0, 0};
    const int NCB1 = ceil_div(a.n, NB), npad = NCB1 * NB;
    const int trank = (TEAM == 1) ? (int)cluster_rank() : (TEAM == 2) ? (int)blockIdx.x : 0;
    const unsigned int team = (TEAM == 1) ? cluster_id_x() : blockIdx.x, nteams = (TEAM == 1) ? cluster_count_x() : gridDim.x;
    double* my_z = zbuf + (size_t)blockIdx.x * slot_z;
    for (unsigned int round = 0;; round++) {
        unsigned int u;
        if constexpr (TEAM == 2) u = round;
        else if constexpr (TEAM == 1) u = team + round * nteams;
        else {
            if (threadIdx.x == 0) job = atomicAdd(counter, 1u);
            __syncthreads();
            u = job;
        }
        if (u >= (unsigned)a.base_n) break;
        const int b = a.base0 + (int)u, r = b % a.R, c = b / a.R;
        const double* rec = a.samples + ((size_t)a.ret_idx[r] * a.n_chains + c) * a.stride;
        fill_ite_spec(a, rec, 0.0, &spec);          // Kp does not depend on doT
        __syncthreads();
        IteGen gen{&spec, sm.exp2tab};
        factor_run<IteGen, TEAM, false, false, false>(gen, NCB1, NCB1, 1, base_L + (size_t)u * a.slot_lo, my_z, sm, pipe, 1 << 30, nullptr, 0,
                                                      nullptr, nullptr, 0, base_linv + (size_t)u * NCB1 * LINV_D);
        if (trank == 0) {
            if (threadIdx.x == 0) base_info[u] = sm.out.info;
            for (int i = threadIdx.x; i < npad; i += blockDim.x) base_z[(size_t)u * npad + i] = my_z[i];
        }
        if constexpr (TEAM == 1) cluster_barrier();
        else if constexpr (TEAM == 2) grid_barrier(gbar);
        else __syncthreads();
    }
}

template <int TEAM, bool PRE = false>
__global__ void __launch_bounds__(FTHREADS, CTAS_PER_SM)
ite_kernel(EstArgs a, double* scratch, size_t slot_scratch, double* zbuf, size_t slot_z, double* xibuf, unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    __shared__ IteSpec spec;
    __shared__ unsigned int job;
    factor_smem_init(sm);
    Pipe pipe{0, 0};
    const int NCB1 = ceil_div(a.n, NB), NCB = 2 * NCB1, npad = NCB1 * NB;
    const int trank = TEAM ? (int)cluster_rank() : 0, tsize = TEAM ? (int)cluster_size() : 1;
    const unsigned int team = TEAM ? cluster_id_x() : blockIdx.x, nteams = TEAM ? cluster_count_x() : gridDim.x;
    double* my_scratch = scratch + (size_t)team * slot_scratch;
    double* my_z = zbuf + (size_t)blockIdx.x * slot_z;
    double* xi = xibuf + (size_t)blockIdx.x * 8 * npad;      // [8 draws][npad columns]
    const unsigned int nbase = (unsigned)a.n_chains * a.R;
    const unsigned int total = PRE ? (unsigned)a.n_doT * a.base_n : (unsigned)a.n_doT * nbase;
    // PRE: the lower-left and lower-right blocks only (block rows >= NCB1) live in this task's scratch
    const size_t hi_off = PRE ? row_off(NCB1) : 0;
    for (unsigned int round = 0;; round++) {
        unsigned int t;
        if constexpr (TEAM != 0) {
            t = team + round * nteams;
        } else {
            if (threadIdx.x == 0) job = atomicAdd(counter, 1u);
            __syncthreads();
            t = job;
        }
        if (t >= total) break;
        unsigned int ub = 0;                 // base task inside this launch's chunk (PRE)
        if constexpr (PRE) { ub = t % a.base_n; t = (t / a.base_n) * nbase + a.base0 + ub; }     // -> global task index d * nbase + b
        const int r = t % a.R, c = (t / a.R) % a.n_chains, d = t / (a.R * a.n_chains);
        const double* rec = a.samples + ((size_t)a.ret_idx[r] * a.n_chains + c) * a.stride;
        fill_ite_spec(a, rec, a.doT[d], &spec);
        if constexpr (PRE)
            for (int i = threadIdx.x; i < npad; i += blockDim.x) my_z[i] = a.base_z[(size_t)ub * npad + i];     // L11^-1 Y of the shared factor
        __syncthreads();
        IteGen gen{&spec, sm.exp2tab};
        double* cov = a.cov_out ? a.cov_out + (size_t)t * a.n * a.n : nullptr;
        if constexpr (PRE)
            factor_run<IteGen, TEAM, true, false, true>(gen, NCB, NCB, 1, my_scratch, my_z, sm, pipe, NCB1, cov, a.n,
                                                        a.base_L + (size_t)ub * a.slot_lo, a.base_linv + (size_t)ub * NCB1 * LINV_D, NCB1);
        else
            factor_run<IteGen, TEAM, true>(gen, NCB, NCB, 1, my_scratch, my_z, sm, pipe, NCB1, cov, a.n);
        int info = sm.out.info;
        if constexpr (PRE) { if (a.base_info[ub] != 0) info = a.base_info[ub]; }
        if (threadIdx.x == 0 && trank == 0 && a.info) a.info[t] = info;
        // MeanITE = -(pre-solve residual of the second block); every CTA of a team holds the same residual in its own zbuf
        const double* wres = my_z + (size_t)MAXRHS * (2 * npad) + npad;
        if (a.mean_out && trank == 0)
            for (int i = threadIdx.x; i < a.n; i += blockDim.x) a.mean_out[(size_t)t * a.n + i] = -wres[i];
        // draws: MeanITE + L22 xi   (src/estimation.jl:95-109; one factor serves all spp draws), eight draws at a time as a tensor-core
        // product: the scratch's atom layout IS the DMMA A-fragment layout (one coalesced 256-byte load per 8 x 4 piece of L22), the B
        // fragment is xi[4 columns][8 draws]. 8-row groups are dealt to the warps (and, in team mode, to the CTAs of the team); the
        // blocks above the diagonal are never touched and the diagonal blocks hold zeros above the diagonal. (The first version walked
        // L22 one row per thread through elem_off: 4.6 ms per task at n = 1024 and 10 draws, latency-bound - 15 % of the task.)
        if (a.ite_out && a.spp > 0) {
            const unsigned gchain = (unsigned)(a.chain0 + c);
            const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
            const int g = lane >> 2, q = lane & 3;
            const int ngroups = (a.n + 7) >> 3;
            for (int s0 = 0; s0 < a.spp; s0 += 8) {
                const int ns = min(8, a.spp - s0);
                __syncthreads();
                for (int e = threadIdx.x; e < 8 * npad; e += blockDim.x) {
                    const int s = e / npad, col = e - s * npad;
                    double v = 0.0;                     // unused draws and the padding columns multiply zeros / identity padding
                    if (s < ns && col < a.n) {
                        Stream st(a.seed, gchain, (uint32_t)(r * a.spp + s0 + s), stream_b(TAG_ITE, (uint32_t)(a.dot0 + d)));
                        v = st.normal_at(col);
                    }
                    xi[(size_t)s * npad + col] = v;
                }
                __syncthreads();
                for (int rg = trank * FWARPS + warp; rg < ngroups; rg += tsize * FWARPS) {
                    const int ib = rg >> 3, r8 = rg & 7;
                    double acc[2] = {0.0, 0.0};
                    const double* xq = xi + (size_t)g * npad + q;          // B fragment: B[k = q][n = g] = xi[draw g][column c0 + q]
                    for (int jb = 0; jb <= ib; jb++) {
                        const double* blk = my_scratch + (block_off(NCB1 + ib, NCB1 + jb, NCB) - hi_off) + r8 * (K4S * 32) + lane;
                        const double* xb = xq + jb * NB;
#pragma unroll 4
                        for (int kk = 0; kk < NB / 4; kk++)                // slab kk / 4, k4 = kk % 4: atoms of one row group are 256 B apart
                            dmma(acc, blk[(kk >> 2) * SLAB_D + (kk & 3) * 32], xb[kk * 4]);
                    }
                    const int row = rg * 8 + g;
                    if (row < a.n) {
                        const double m = -wres[row];
                        double* o = a.ite_out + (((size_t)d * a.n_chains + c) * a.R * a.spp + (size_t)r * a.spp + s0) * a.n + row;
                        if (2 * q < ns) o[(size_t)(2 * q) * a.n] = m + acc[0];
                        if (2 * q + 1 < ns) o[(size_t)(2 * q + 1) * a.n] = m + acc[1];
                    }
                }
            }
        }
        // the next task overwrites the shared scratch: wait until every CTA of the team has finished reading L22
        if constexpr (TEAM != 0) cluster_barrier();
        else __syncthreads();
    }
}

// SATE fast path: one task = (doT, chain, retained sample)
__global__ void __launch_bounds__(FTHREADS, CTAS_PER_SM)
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
            rspec.feat[spec.D] = a.T; rspec.w[spec.D] = spec.wT; rspec.sw[spec.D] = sqrt(spec.wT);
        }
        for (int k = threadIdx.x; k < spec.D; k += blockDim.x) { rspec.feat[k] = spec.feat[k]; rspec.w[k] = spec.w[k]; rspec.sw[k] = sqrt(spec.w[k]); }
        __syncthreads();
        RbfGen rgen{&rspec, sm.exp2tab};
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
                    Stream st(a.seed, gchain, (uint32_t)(r * a.spp + s), stream_b(TAG_SATE, (uint32_t)(a.dot0 + d)));
                    a.sate_out[((size_t)d * a.n_chains + c) * a.R * a.spp + (size_t)r * a.spp + s] = ms + sd * st.normal();
                }
            }
        }
        __syncthreads();
    }
}

// launch one of the ite_kernel instantiations on `grid` teams of `team` CTAs
template <bool PRE>
static int launch_ite_kernel(Ctx* ctx, const EstArgs& a, int team, int grid, double* xi) {
    if (team == 1) {
        GP_CUDA(ctx, cudaFuncSetAttribute(ite_kernel<0, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
        ite_kernel<0, PRE><<<grid, FTHREADS, sizeof(FactorSmem), ctx->stream>>>(a, ctx->scratch, ctx->slot_scratch_d, ctx->zbuf, ctx->slot_z_d, xi,
                                                                                ctx->counter);
        ctx->launches++;
        GP_CUDA(ctx, cudaGetLastError());
        return GPSLC_OK;
    }
    GP_CUDA(ctx, cudaFuncSetAttribute(ite_kernel<1, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
    if (team > 8) GP_CUDA(ctx, cudaFuncSetAttribute(ite_kernel<1, PRE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = team; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(FTHREADS); cfg.dynamicSmemBytes = sizeof(FactorSmem); cfg.stream = ctx->stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(grid * team);
    cudaError_t e = cudaLaunchKernelEx(&cfg, ite_kernel<1, PRE>, a, ctx->scratch, ctx->slot_scratch_d, ctx->zbuf, ctx->slot_z_d, xi, ctx->counter);
    ctx->launches++;
    if (e != cudaSuccess) return ctx->cuda_fail(e, "ite_kernel");
    return GPSLC_OK;
}

// number of teams of `team` CTAs that can be resident for kernel k (clusters), clipped to `grid`
template <class K>
static int clip_to_resident_clusters(Ctx* ctx, K k, int team, int grid, int* out) {
    GP_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
    if (team > 8) GP_CUDA(ctx, cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = team; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(FTHREADS); cfg.dynamicSmemBytes = sizeof(FactorSmem); cfg.stream = ctx->stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(grid * team);
    int max_clusters = 0;
    GP_CUDA(ctx, cudaOccupancyMaxActiveClusters(&max_clusters, k, &cfg));
    if (max_clusters < 1) return ctx->fail(GPSLC_ERR_CUDA, "no resident cluster of the requested size");
    *out = grid > max_clusters ? max_clusters : grid;
    return GPSLC_OK;
}

// Several doT values per posterior sample (predictCounterfactualEffects, src/prediction.jl:23-36): chol(Kp), L11^-1 Y and the panels'
// P2 outputs are computed ONCE per (chain, retained sample) by ite_base_kernel and every (doT, chain, sample) task continues from
// them (factor.cuh, PRE): 7 n^3 / 3 flops per doT instead of 8 n^3 / 3. The base tasks are processed in chunks that fit a fraction
// of the free device memory. GPSLC_ITE_SHARE=0 forces the fused one-Cholesky-per-task path (development knob).
static int launch_ite_shared(Ctx* ctx, const EstArgs& a0) {
    EstArgs a = a0;
    const int NCB1 = ceil_div(a.n, NB), NCB = 2 * NCB1, npad = NCB1 * NB;
    const long long nbase = (long long)a.n_chains * a.R;
    const size_t slot_lo = row_off(NCB1);
    const size_t per_base = (slot_lo + (size_t)NCB1 * LINV_D + npad) * sizeof(double) + sizeof(int);
    size_t free_b = 0, total_b = 0;
    GP_CUDA(ctx, cudaMemGetInfo(&free_b, &total_b));
    long long chunk = (long long)((double)free_b * 0.25 / (double)per_base);
    if (chunk < 1) return ctx->fail(GPSLC_ERR_CUDA, "not enough device memory for one shared Kp factor");
    if (chunk > nbase) chunk = nbase;
    double *bL = nullptr, *blinv = nullptr, *bz = nullptr; int* binfo = nullptr; unsigned int* gbar = nullptr;
    GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&bL), (size_t)chunk * slot_lo * sizeof(double)));
    GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&blinv), (size_t)chunk * NCB1 * LINV_D * sizeof(double)));
    GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&bz), (size_t)chunk * npad * sizeof(double)));
    GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&binfo), (size_t)chunk * sizeof(int)));
    GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&gbar), 2 * sizeof(unsigned int)));
    double* xi = nullptr;
    size_t xi_cap = 0;
    for (long long b0 = 0; b0 < nbase; b0 += chunk) {
        const int nb = (int)((nbase - b0 < chunk) ? nbase - b0 : chunk);
        a.base0 = (int)b0; a.base_n = nb; a.base_L = bL; a.base_linv = blinv; a.base_z = bz; a.base_info = binfo; a.slot_lo = slot_lo;
        // ---- phase A: the shared factors of this chunk (they go to bL; only the per-CTA z buffers of the context are used)
        {
            const int gmax = ctx->num_sms < NCB1 ? ctx->num_sms : NCB1;        // grid mode: at most one CTA per block row, one per SM
            if ((long long)nb * 16 <= gmax) {
                // few large samples (the n = 8192 sweep has ONE): the whole cooperatively launched grid factors one Kp after the other
                GP_TRY(ensure_zbuf(ctx, NCB1, gmax));
                GP_CUDA(ctx, cudaMemsetAsync(gbar, 0, 2 * sizeof(unsigned int), ctx->stream));
                GP_CUDA(ctx, cudaFuncSetAttribute(ite_base_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
                void* args[] = {(void*)&a, (void*)&bL, (void*)&blinv, (void*)&bz, (void*)&binfo, (void*)&ctx->zbuf, (void*)&ctx->slot_z_d,
                                (void*)&ctx->counter, (void*)&gbar};
                cudaError_t e = cudaLaunchCooperativeKernel((void*)ite_base_kernel<2>, dim3(gmax), dim3(FTHREADS), args, sizeof(FactorSmem), ctx->stream);
                ctx->launches++;
                if (e != cudaSuccess) return ctx->cuda_fail(e, "ite_base_kernel (cooperative)");
            } else {
                const int team = pick_team(ctx, nb, NCB1);
                const int per = (team > 1 && NCB1 >= 32) ? 1 : 2;
                long long grid = (long long)per * ctx->num_sms / team;
                if (grid > nb) grid = nb;
                int g = (int)grid;
                if (team > 1) GP_TRY(clip_to_resident_clusters(ctx, ite_base_kernel<1>, team, g, &g));
                GP_TRY(ensure_zbuf(ctx, NCB1, (long long)g * team));
                GP_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
                if (team > 1) {
                    cudaLaunchConfig_t cfg = {};
                    cudaLaunchAttribute attr[1];
                    attr[0].id = cudaLaunchAttributeClusterDimension;
                    attr[0].val.clusterDim.x = team; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
                    cfg.blockDim = dim3(FTHREADS); cfg.dynamicSmemBytes = sizeof(FactorSmem); cfg.stream = ctx->stream;
                    cfg.attrs = attr; cfg.numAttrs = 1;
                    cfg.gridDim = dim3(g * team);
                    cudaError_t e = cudaLaunchKernelEx(&cfg, ite_base_kernel<1>, a, bL, blinv, bz, binfo, ctx->zbuf, ctx->slot_z_d, ctx->counter, gbar);
                    ctx->launches++;
                    if (e != cudaSuccess) return ctx->cuda_fail(e, "ite_base_kernel");
                } else {
                    GP_CUDA(ctx, cudaFuncSetAttribute(ite_base_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
                    ite_base_kernel<0><<<g, FTHREADS, sizeof(FactorSmem), ctx->stream>>>(a, bL, blinv, bz, binfo, ctx->zbuf, ctx->slot_z_d, ctx->counter, gbar);
                    ctx->launches++;
                    GP_CUDA(ctx, cudaGetLastError());
                }
            }
        }
        // ---- phase B: every (doT, base task) of the chunk continues from its shared factor
        {
            const long long total = (long long)a.n_doT * nb;
            int team = pick_team(ctx, total, NCB);
            if (const char* e = getenv("GPSLC_ITE_TEAM")) { const int g = atoi(e); if (g >= 1 && g <= 16) team = g; }   // development knob; > 8 = non-portable cluster
            int grid = 0;
            GP_TRY(ensure_workspace(ctx, NCB, NCB, total, &grid, team, NCB1));
            if (team > 1) GP_TRY(clip_to_resident_clusters(ctx, ite_kernel<1, true>, team, grid, &grid));
            GP_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
            const size_t need_xi = (size_t)grid * team * 8 * npad * sizeof(double);
            if (need_xi > xi_cap) { GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&xi), need_xi)); xi_cap = need_xi; }
            GP_TRY(launch_ite_kernel<true>(ctx, a, team, grid, xi));
        }
    }
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    if (e2 != cudaSuccess) return ctx->cuda_fail(e2, "ite_kernel (shared Kp)");
    return GPSLC_OK;
}

int launch_ite(Ctx* ctx, const EstArgs& a) {
    const int NCB = 2 * ceil_div(a.n, NB);
    const long long total = (long long)a.n_doT * a.n_chains * a.R;
    if (total == 0) return GPSLC_OK;
    {
        const char* e = getenv("GPSLC_ITE_SHARE");
        const bool share = e ? (atoi(e) != 0) : true;
        if (share && a.n_doT >= 2) return launch_ite_shared(ctx, a);
    }
    int team = pick_team(ctx, total, NCB);
    int grid = 0;   // number of scratch slots == number of teams
    GP_TRY(ensure_workspace(ctx, NCB, NCB, total, &grid, team));
    if (team > 1) GP_TRY(clip_to_resident_clusters(ctx, ite_kernel<1, false>, team, grid, &grid));
    GP_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
    double* xi = nullptr;
    GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&xi), (size_t)grid * team * 8 * (NCB / 2) * NB * sizeof(double)));
    int rc = launch_ite_kernel<false>(ctx, a, team, grid, xi);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    if (rc) return rc;
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
    GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&d1), (size_t)grid * a.n * sizeof(double)));
    sate_kernel<<<grid, FTHREADS, sizeof(FactorSmem), ctx->stream>>>(a, ctx->scratch, ctx->slot_scratch_d, ctx->zbuf, ctx->slot_z_d, d1, ctx->counter);
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    if (e != cudaSuccess) return ctx->cuda_fail(e, "sate_kernel");
    if (e2 != cudaSuccess) return ctx->cuda_fail(e2, "sate_kernel");
    return GPSLC_OK;
}

// ------------------------------------------------------------------------------------------------ summarizeEstimates
// One CTA per (batch element, tile of IT individuals): the tile's samples are staged through shared memory with coalesced
// loads (individual fastest), each warp then bitonic-sorts the rows it owns and lane 0 interpolates the two quantiles
// (src/driver.jl:129-149; Julia `quantile` default = type 7). HBM-bound: every sample is read once (8 m n bytes per batch
// element), 24 n bytes written.
__global__ void __launch_bounds__(256) summarize_kernel(const double* __restrict__ samples, int m, int mpad, int n, int IT, double lowerQ,
                                                        double upperQ, double* __restrict__ out) {
    extern __shared__ double tile[];   // [IT][mpad]
    const int b = blockIdx.y, i0 = blockIdx.x * IT;
    const double* src = samples + (size_t)b * m * n;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    for (int e = threadIdx.x; e < IT * mpad; e += blockDim.x) {
        const int s = e / IT, ii = e - s * IT;
        tile[ii * mpad + s] = (s < m && i0 + ii < n) ? src[(size_t)s * n + i0 + ii] : INFINITY;
    }
    __syncthreads();
    for (int ii = warp; ii < IT && i0 + ii < n; ii += nwarps) {
        double* x = tile + ii * mpad;
        for (int k = 2; k <= mpad; k <<= 1) {
            for (int jj = k >> 1; jj > 0; jj >>= 1) {
                for (int t = lane; t < (mpad >> 1); t += 32) {
                    const int lo = ((t & ~(jj - 1)) << 1) | (t & (jj - 1));   // element index with bit jj cleared
                    const int hi = lo | jj;
                    const bool up = ((lo & k) == 0);
                    const double a = x[lo], c = x[hi];
                    if ((a > c) == up) { x[lo] = c; x[hi] = a; }
                }
                __syncwarp();
            }
        }
        double sum = 0.0;
        for (int t = lane; t < m; t += 32) sum += x[t];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            double* o3 = out + ((size_t)b * n + i0 + ii) * 3;
            o3[0] = sum / m;
            const double qs[2] = {lowerQ, upperQ};
            for (int qi = 0; qi < 2; qi++) {
                const double h = (m - 1) * qs[qi];
                int l = (int)floor(h);
                l = max(0, min(l, m - 1));
                const int u = min(l + 1, m - 1);
                o3[1 + qi] = x[l] + (h - l) * (x[u] - x[l]);
            }
        }
    }
}

// Large sample counts (m > 8192: e.g. the draws of all 512 chains pooled, 512 x 150 = 76800 per individual) do not fit a shared-memory
// sort. The same three statistics by radix selection instead: one CTA per group of 4 adjacent individuals (one 32-byte sector per
// sample), eight passes over the samples, one per 8-bit digit of the order-preserving 64-bit key, most significant first; every pass
// builds the digit histograms of all 16 (individual, order statistic) selections at once among the samples that match the selection's
// prefix so far. The two quantiles need the order statistics l and l+1 each (type 7). HBM-bound: 8 reads of the samples.
__device__ __forceinline__ unsigned long long f64_key(double x) {
    const unsigned long long b = (unsigned long long)__double_as_longlong(x);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__device__ __forceinline__ double key_f64(unsigned long long k) {
    const unsigned long long b = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    return __longlong_as_double((long long)b);
}

__global__ void __launch_bounds__(256) summarize_select_kernel(const double* __restrict__ samples, int m, int n, double lowerQ, double upperQ,
                                                               double* __restrict__ out) {
    __shared__ unsigned int hist[16][256];
    __shared__ unsigned long long prefix[16];
    __shared__ unsigned int rank[16];
    __shared__ double red[8][4];
    const int b = blockIdx.y, i0 = blockIdx.x * 4;
    const int ni = min(4, n - i0);
    const double* src = samples + (size_t)b * m * n + i0;
    const int tid = threadIdx.x;
    if (tid < 16) {
        const int which = tid & 3;                      // order statistic: lower l, lower l+1, upper l, upper l+1
        const double h = (m - 1) * ((which < 2) ? lowerQ : upperQ);
        int l = (int)floor(h);
        l = max(0, min(l, m - 1));
        rank[tid] = (unsigned)((which & 1) ? min(l + 1, m - 1) : l);
        prefix[tid] = 0ull;
    }
    // mean
    double sum[4] = {0.0, 0.0, 0.0, 0.0};
    for (int s = tid; s < m; s += blockDim.x)
        for (int k = 0; k < ni; k++) sum[k] += src[(size_t)s * n + k];
    for (int k = 0; k < 4; k++) {
        double v = sum[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((tid & 31) == 0) red[tid >> 5][k] = v;
    }
    __syncthreads();
    for (int pass = 0; pass < 8; pass++) {
        const int shift = 56 - 8 * pass;
        for (int e = tid; e < 16 * 256; e += blockDim.x) (&hist[0][0])[e] = 0u;
        __syncthreads();
        for (int s = tid; s < m; s += blockDim.x) {
            for (int k = 0; k < ni; k++) {
                const unsigned long long key = f64_key(src[(size_t)s * n + k]);
                const unsigned int digit = (unsigned int)(key >> shift) & 255u;
                const unsigned long long hi = (pass == 0) ? 0ull : (key >> (shift + 8));
#pragma unroll
                for (int w = 0; w < 4; w++)
                    if (pass == 0 || hi == prefix[k * 4 + w]) atomicAdd(&hist[k * 4 + w][digit], 1u);
            }
        }
        __syncthreads();
        if (tid < 16 && (tid >> 2) < ni) {
            unsigned int r = rank[tid], d = 0;
            while (d < 255u && r >= hist[tid][d]) { r -= hist[tid][d]; d++; }
            rank[tid] = r;
            prefix[tid] = (prefix[tid] << 8) | d;
        }
        __syncthreads();
    }
    if (tid < ni) {
        double* o3 = out + ((size_t)b * n + i0 + tid) * 3;
        double tot = 0.0;
        for (int wv = 0; wv < 8; wv++) tot += red[wv][tid];
        o3[0] = tot / m;
        const double qs[2] = {lowerQ, upperQ};
        for (int qi = 0; qi < 2; qi++) {
            const double h = (m - 1) * qs[qi];
            int l = (int)floor(h);
            l = max(0, min(l, m - 1));
            const double xl = key_f64(prefix[tid * 4 + 2 * qi]), xu = key_f64(prefix[tid * 4 + 2 * qi + 1]);
            o3[1 + qi] = xl + (h - l) * (xu - xl);
        }
    }
}

// Subgroup average of the docs example (docs/src/index.md:101-108: `mean(ite[:, maIdx, :], dims = 2)`): one warp per row
// (batch element b, draw s) of samples [batch][m][n] sums the masked individuals — coalesced, every sample read once (HBM-bound,
// 8 n bytes per row), fixed summation order (lane-strided partial sums, then a shuffle tree). groups > 0 transposes the result for the
// summary that follows: b = d * groups + c goes to out[c][s][d] (per chain the column-major nDoT x nSamples matrix the docs example
// hands to summarizeEstimates); groups == 0: out[b][s].
__global__ void __launch_bounds__(256) subset_mean_kernel(const double* __restrict__ samples, const unsigned char* __restrict__ mask,
                                                          size_t rows, int m, int n, int groups, int n_d, double inv_count,
                                                          double* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const size_t wpg = (size_t)gridDim.x * (blockDim.x >> 5);
    for (size_t row = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); row < rows; row += wpg) {
        const double* src = samples + row * n;
        double sum = 0.0;
        for (int i = lane; i < n; i += 32)
            if (mask[i]) sum += src[i];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) {
            size_t oi = row;
            if (groups > 0) {
                const size_t b = row / m, s = row - b * m, d = b / groups, c = b - d * groups;
                oi = (c * m + s) * n_d + d;
            }
            out[oi] = sum * inv_count;
        }
    }
}

int launch_subset_mean(Ctx* ctx, const double* samples, const unsigned char* mask, size_t rows, int m, int n, int groups, int n_d,
                       int count, double* out) {
    const size_t want = (rows + 7) / 8;
    const int grid = (int)(want < (size_t)(148 * 8) ? want : (size_t)(148 * 8));
    subset_mean_kernel<<<grid, 256, 0, ctx->stream>>>(samples, mask, rows, m, n, groups, n_d, 1.0 / count, out);
    ctx->launches++;
    GP_CUDA(ctx, cudaGetLastError());
    return GPSLC_OK;
}

int launch_summarize(Ctx* ctx, const double* samples, int batch, int m, int n, double ci, double* out) {
    int mpad = 2;
    while (mpad < m) mpad <<= 1;
    const double lowerQ = (1.0 - ci) / 2.0, upperQ = 1.0 - lowerQ;
    if (mpad > 8192) {
        dim3 grid(ceil_div(n, 4), batch);
        summarize_select_kernel<<<grid, 256, 0, ctx->stream>>>(samples, m, n, lowerQ, upperQ, out);
        ctx->launches++;
        GP_CUDA(ctx, cudaGetLastError());
        return GPSLC_OK;
    }
    int IT = (int)((192 * 1024) / ((size_t)mpad * sizeof(double)));
    if (IT > 32) IT = 32;
    if (IT > n) IT = n;
    const size_t smem = (size_t)IT * mpad * sizeof(double);
    GP_CUDA(ctx, cudaFuncSetAttribute(summarize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid(ceil_div(n, IT), batch);
    summarize_kernel<<<grid, 256, smem, ctx->stream>>>(samples, m, mpad, n, IT, lowerQ, upperQ, out);
    ctx->launches++;
    GP_CUDA(ctx, cudaGetLastError());
    return GPSLC_OK;
}

}  // namespace gpslc
