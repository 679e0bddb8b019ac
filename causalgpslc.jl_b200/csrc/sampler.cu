// Many-chain posterior sampler: Gen `generate` + nOuter x (nMHInner single-site MH sweeps + nESInner elliptical-slice
// passes), i.e. all eight `Posterior` methods of src/inference.jl:4-379, for thousands of independent chains.
//
// What differs from the reference by design (values identical, cost not — SURVEY.md App. A4/B2):
//  * a single-site update re-scores only the GP factor the site feeds (the reference re-executes the whole model);
//  * within a sweep the sites of different factors are conditionally independent given U, so each (chain, factor)
//    "lane" runs its own sites sequentially on one CTA while other lanes run elsewhere; with the counter-based RNG
//    keyed by (chain, sweep, site) the result is the one the sequential sweep would produce;
//  * the U prior N(0, uNoise*SigmaU) uses the closed form of the block matrix — no factorisation.
#include <algorithm>
#include <string>
#include "sampler.cuh"
#include "capi_util.cuh"

namespace gpslc {

// ------------------------------------------------------------------------------------------------ device helpers

// not inlined: log + lgamma at four call sites were 1500 instructions of the sampler kernels
__device__ __noinline__ double ig_logpdf(double x, double shape, double scale) {
    if (!(x > 0.0)) return -INFINITY;
    return shape * log(scale) - lgamma(shape) - (shape + 1.0) * log(x) - scale / x;
}

__device__ inline double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double s = 0.0;
    for (int w = 0; w < nw; w++) s += red[w];
    return s;
}

// u' SigmaU^-1 u for the block matrix of src/utils.jl:17-33 (each block cov*11' + d*I), whole CTA cooperates
__device__ inline double block_u_quad(const ModelDev& m, const double* u, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double part = 0.0;
    for (int o = warp; o < m.n_obj; o += nw) {
        const int s0 = m.obj_start[o], cnt = m.obj_start[o + 1] - s0;
        double s = 0.0;
        for (int i = lane; i < cnt; i += 32) s += u[s0 + i];
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
        const double mean = s / cnt;
        double dv = 0.0;
        for (int i = lane; i < cnt; i += 32) { const double t = u[s0 + i] - mean; dv = fma(t, t, dv); }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) dv += __shfl_xor_sync(0xffffffffu, dv, off);
        if (lane == 0) part += dv / m.dU + cnt * mean * mean / (m.dU + cnt * m.cov);
    }
    return block_sum(part, red);
}

// ---- dense (unstructured) SigmaU: `samplePosterior(hyperparams, priorparams, SigmaU, X, T, Y)` accepts any matrix
// (src/driver.jl:59-69) and generateU factors uNoise*SigmaU (src/model_prior.jl:27-30). Here L_S = chol(SigmaU) is computed once per
// sampler (m.SL, factor.cuh scratch layout, identity-padded to a multiple of 64); chol(uNoise*SigmaU) = sqrt(uNoise) L_S.
struct USmem { double red[32]; double wv[NB]; };

// u' SigmaU^-1 u = |L_S^-1 u|^2 by blocked forward substitution; whole CTA of 256 threads; zs: [npad] scratch of this chain
__device__ inline double dense_u_quad(const ModelDev& m, const double* u, double* zs, USmem& us) {
    const int NCB = ceil_div(m.n, NB);
    const int tid = threadIdx.x, r = tid >> 2, kq = tid & 3;
    __syncthreads();
    for (int jb = 0; jb < NCB; jb++) {
        double s = 0.0;
        for (int J = 0; J < jb; J++) {
            const double* blk = m.SL + block_off(jb, J, NCB);
            const double* zj = zs + J * NB;
            for (int cc = kq; cc < NB; cc += 4) s = fma(blk[elem_off(r, cc)], zj[cc], s);
        }
        s += __shfl_xor_sync(0xffffffffu, s, 1);
        s += __shfl_xor_sync(0xffffffffu, s, 2);
        const int gi = jb * NB + r;
        if (kq == 0) us.wv[r] = ((gi < m.n) ? u[gi] : 0.0) - s;
        __syncthreads();
        if (tid < 32) {
            // rows `lane` and `lane + 32` of the 64 x 64 lower-triangular diagonal block
            const double* blk = m.SL + block_off(jb, jb, NCB);
            double w0 = us.wv[tid], w1 = us.wv[tid + 32];
            for (int cc = 0; cc < NB; cc++) {
                const double wc = __shfl_sync(0xffffffffu, (cc < 32) ? w0 : w1, cc & 31);
                const double zc = wc / blk[elem_off(cc, cc)];
                if (tid == (cc & 31)) zs[jb * NB + cc] = zc;
                if (tid > cc) w0 = fma(-blk[elem_off(tid, cc)], zc, w0);
                if (tid + 32 > cc) w1 = fma(-blk[elem_off(tid + 32, cc)], zc, w1);
            }
        }
        __syncthreads();
    }
    double part = 0.0;
    for (int i = tid; i < m.n; i += blockDim.x) part = fma(zs[i], zs[i], part);
    return block_sum(part, us.red);
}

__device__ inline double u_quad(const ModelDev& m, const ChainDev& c, int chain, const double* u, USmem& us) {
    if (m.dense_u) return dense_u_quad(m, u, c.zs + (size_t)chain * ceil_div(m.n, NB) * NB, us);
    return block_u_quad(m, u, us.red);
}

// out ~ N(0, uNoise*SigmaU) from the normals of stream `st`: block structure sqrt(uNoise) (sqrt(cov) z_obj + sqrt(d) z_i) (exact for the
// matrix of src/utils.jl:17-33), dense sqrt(uNoise) L_S z. Whole CTA.
__device__ inline void u_prior_draw(const ModelDev& m, const ChainDev& c, int chain, const Stream& st, double un, double* out) {
    const double su = sqrt(un);
    if (m.dense_u) {
        double* zs = c.zs + (size_t)chain * ceil_div(m.n, NB) * NB;
        __syncthreads();
        for (int i = threadIdx.x; i < m.n; i += blockDim.x) zs[i] = st.normal_at(i);
        __syncthreads();
        for (int i = threadIdx.x; i < m.n; i += blockDim.x) {
            double accv[1];
            tri_matvec_row<1>(m.SL, ceil_div(m.n, NB), 0, 0, i, m.n, zs, 0, 1, accv);
            out[i] = su * accv[0];
        }
        __syncthreads();
    } else {
        const double sc = sqrt(m.cov), sd = sqrt(m.dU);
        for (int i = threadIdx.x; i < m.n; i += blockDim.x)
            out[i] = su * (sc * st.normal_at(m.n + m.obj_of[i]) + sd * st.normal_at(i));
    }
}

// element (r, c) of the n x nU matrix the model's kernels see (src/model_likelihood.jl:7 + src/utils.jl:60-64)
__device__ __forceinline__ void ueff_src(const ModelDev& m, int r, int c, int& a, int& b) {
    if (m.u_layout_reference) { const long long f = (long long)r + (long long)c * m.n; a = (int)(f % m.nU); b = (int)(f / m.nU); }
    else { a = c; b = r; }
}

// rebuild Ueff (or UeffP with U_k replaced by `repl`) for one chain
__device__ inline void build_ueff(const ModelDev& m, const double* U, int k, const double* repl, double* out) {
    const int total = m.n * m.nU;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int c = i / m.n, r = i - c * m.n;
        int a, b;
        ueff_src(m, r, c, a, b);
        out[i] = (repl && a == k) ? repl[b] : U[(size_t)a * m.n + b];
    }
}

__device__ inline void build_spec(const ModelDev& m, const ChainDev& c, int chain, int f, const double* Ubase, int ov_param,
                                  double ov_val, RbfSpec* spec) {
    const FactorDef& fd = m.fdef[f];
    const double* theta = c.theta + (size_t)chain * m.n_params;
    const double* Xd = m.X + (size_t)chain * m.xstride;     // this chain's observed data (per-chain datasets: SBC-style runs)
    const double* Td = m.T + (size_t)chain * m.tstride;
    const double* Yd = m.Y + (size_t)chain * m.tstride;
    const double* Xm = m.has_xmodel ? c.Xmodel + (size_t)chain * m.n * m.nX : Xd;
    for (int d = threadIdx.x; d < fd.D; d += blockDim.x) {
        const int kind = fd.src_kind[d], idx = fd.src_idx[d];
        spec->feat[d] = (kind == SRC_U) ? Ubase + (size_t)idx * m.n : (kind == SRC_X) ? Xm + (size_t)idx * m.n : Td;
        const int p = fd.ls_param[d];
        const double ls = (p == ov_param) ? ov_val : theta[p];
        spec->w[d] = m.ls_unsquared ? 1.0 / ls : 1.0 / (ls * ls);
        spec->sw[d] = m.ls_unsquared ? rsqrt(ls) : 1.0 / ls;
    }
    if (threadIdx.x == 0) {
        spec->D = fd.D;
        spec->n = m.n;
        spec->scale = (fd.scale_param == ov_param) ? ov_val : theta[fd.scale_param];
        spec->noise = (fd.noise_param == ov_param) ? ov_val : theta[fd.noise_param];
        const double* y = (fd.target_kind == TGT_XCOL) ? Xd + (size_t)fd.target_idx * m.n
                          : (fd.target_kind == TGT_T)  ? Td
                          : (fd.target_kind == TGT_LOGIT) ? c.logitT + (size_t)chain * m.n : Yd;
        spec->y[0] = y; spec->y[1] = y;
    }
}

__device__ __forceinline__ double logpdf_from(const FactorOut& o, int n) {
    return (o.info == 0) ? -0.5 * (n * LOG_2PI + o.logdet + o.gram[0]) : -INFINITY;
}

// ------------------------------------------------------------------------------------------------ init (`generate`)
__global__ void __launch_bounds__(256) init_chains_kernel(ModelDev m, ChainDev c) {
    __shared__ USmem us;
    const int chain = blockIdx.x;
    const unsigned gchain = (unsigned)(m.chain0 + chain);
    double* theta = c.theta + (size_t)chain * m.n_params;
    for (int p = threadIdx.x; p < m.n_params; p += blockDim.x) theta[p] = nan("");
    __syncthreads();
    for (int s = threadIdx.x; s < m.n_sites; s += blockDim.x) {
        const SiteDef sd = m.sites[s];
        Stream st(m.seed, gchain, (uint32_t)sd.param, stream_b(TAG_INIT_PARAM, 0));
        theta[sd.param] = st.inv_gamma(sd.pshape, sd.pscale);
    }
    __syncthreads();
    for (int k = 0; k < m.nU; k++) {
        const double un = theta[0];
        Stream st(m.seed, gchain, (uint32_t)k, stream_b(TAG_INIT_VEC, 0));
        u_prior_draw(m, c, chain, st, un, c.U + ((size_t)chain * m.nU + k) * m.n);
    }
    if (m.binary) {
        // logitT ~ N(0, I) when T has no parents (src/model_prior.jl:194-200); with parents init_logit_kernel overwrites it by L z
        Stream st(m.seed, gchain, (uint32_t)m.nU, stream_b(TAG_INIT_VEC, 0));
        double* lt = c.logitT + (size_t)chain * m.n;
        for (int i = threadIdx.x; i < m.n; i += blockDim.x) lt[i] = st.normal_at(i);
    }
    if (m.has_xmodel) {
        for (int k = 0; k < m.nX; k++) {
            Stream st(m.seed, gchain, (uint32_t)k, stream_b(TAG_INIT_XMODEL, 0));
            double* x = c.Xmodel + ((size_t)chain * m.nX + k) * m.n;
            for (int i = threadIdx.x; i < m.n; i += blockDim.x) x[i] = st.normal_at(i);
        }
    }
    __syncthreads();
    if (m.nU > 0) {
        build_ueff(m, c.U + (size_t)chain * m.nU * m.n, -1, nullptr, c.Ueff + (size_t)chain * m.nU * m.n);
        for (int k = 0; k < m.nU; k++) {
            const double qv = u_quad(m, c, chain, c.U + ((size_t)chain * m.nU + k) * m.n, us);
            if (threadIdx.x == 0) c.q[(size_t)chain * m.nU + k] = qv;
        }
    }
    if (threadIdx.x == 0) c.info[chain] = 0;
}

// recompute the derived per-chain quantities (Ueff, q) after the host overwrote theta/U (gpslc_sampler_set_state)
__global__ void __launch_bounds__(256) refresh_chains_kernel(ModelDev m, ChainDev c) {
    __shared__ USmem us;
    const int chain = blockIdx.x;
    if (m.nU > 0) {
        build_ueff(m, c.U + (size_t)chain * m.nU * m.n, -1, nullptr, c.Ueff + (size_t)chain * m.nU * m.n);
        for (int k = 0; k < m.nU; k++) {
            const double qv = u_quad(m, c, chain, c.U + ((size_t)chain * m.nU + k) * m.n, us);
            if (threadIdx.x == 0) c.q[(size_t)chain * m.nU + k] = qv;
        }
    }
    if (threadIdx.x == 0) c.info[chain] = 0;
}

// L_S = chol(SigmaU) for a dense SigmaU, once per sampler (one CTA: a start-up cost, not part of any sweep)
__global__ void __launch_bounds__(FTHREADS, CTAS_PER_SM) sigma_u_factor_kernel(const double* S, int n, double* SL, double* zbuf, int* info_out) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    factor_smem_init(sm);
    Pipe pipe{0, 0};
    const int NCB = ceil_div(n, NB);
    DenseGen gen{S, {nullptr, nullptr}, n, n};
    factor_run(gen, NCB, NCB, 0, SL, zbuf, sm, pipe);
    if (threadIdx.x == 0) *info_out = sm.out.info;
}

// ------------------------------------------------------------------------------------------------ factor evaluation
// tasks = (chain from list or all chains) x (existing factors); writes lp / lpP
template <int TEAM>
__global__ void __launch_bounds__(FTHREADS, CTAS_PER_SM)
eval_factors_kernel(ModelDev m, ChainDev c, const int* list, const unsigned int* n_list_dev, int n_all, const int* exist,
                    int n_exist, int proposed, double* scratch, size_t slot_scratch, double* zbuf, size_t slot_z,
                    unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    __shared__ RbfSpec spec;
    __shared__ unsigned int job;
    factor_smem_init(sm);
    Pipe pipe{0, 0};
    const int NCB = ceil_div(m.n, NB);
    // TEAM = 1: one thread-block cluster per task (few large tasks), see factor.cuh; tasks dealt round-robin to the clusters
    const int trank = TEAM ? (int)cluster_rank() : 0;
    const unsigned int team = TEAM ? cluster_id_x() : blockIdx.x, nteams = TEAM ? cluster_count_x() : gridDim.x;
    double* my_scratch = scratch + (size_t)team * slot_scratch;
    double* my_z = zbuf + (size_t)blockIdx.x * slot_z;
    const unsigned int n_list = list ? *n_list_dev : (unsigned)n_all;
    const unsigned int total = n_list * (unsigned)n_exist;
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
        const int li = t / n_exist, f = exist[t - li * n_exist];
        const int chain = list ? list[li] : li;
        const double* Ubase = (proposed ? c.UeffP : c.Ueff) + (size_t)chain * m.nU * m.n;
        build_spec(m, c, chain, f, Ubase, -1, 0.0, &spec);
        __syncthreads();
        RbfGen gen{&spec, sm.exp2tab};
        factor_run<RbfGen, TEAM>(gen, NCB, NCB, 1, my_scratch, my_z, sm, pipe, 1 << 30, nullptr, 0, nullptr, nullptr, 0, nullptr, /*keep_diag=*/false);
        if (threadIdx.x == 0 && trank == 0) {
            const FactorOut o = sm.out;
            (proposed ? c.lpP : c.lp)[(size_t)chain * m.nF + f] = logpdf_from(o, m.n);
            if (o.info != 0) {
                if (proposed) atomicOr(&c.infoP[chain], 1);
                else atomicMax(&c.info[chain], o.info);
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ MH lanes
// One task = (chain, lane): all single-site MH updates (src/proposal.jl:32-41 + Gen `mh`) of the
// sites feeding one GP factor, for sweeps [j0, j1) of outer iteration `outer`.
template <int TEAM>
__global__ void __launch_bounds__(FTHREADS, CTAS_PER_SM)
mh_lanes_kernel(ModelDev m, ChainDev c, const int* lane_order, int outer, int j0, int j1, double* scratch, size_t slot_scratch,
                double* zbuf, size_t slot_z, unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    __shared__ RbfSpec spec;
    __shared__ unsigned int job;
    __shared__ double s_new, s_terms[5];
    factor_smem_init(sm);
#ifdef GPSLC_PHASE_TIMING
    const long long _k0 = clock64();
#endif
    Pipe pipe{0, 0};
    const int NCB = ceil_div(m.n, NB);
    // TEAM = 1 (few chains at large n, e.g. the reference's single-chain gpslc() call): one thread-block cluster per (chain, lane)
    // task. Every CTA of the team runs the whole site loop redundantly (same Philox streams, same factor results, hence the
    // same accept decisions); the factorisations are shared (factor.cuh), the chain state is written by rank 0 only and
    // published to the team by the cluster barrier that ends each update.
    const int trank = TEAM ? (int)cluster_rank() : 0;
    const unsigned int team = TEAM ? cluster_id_x() : blockIdx.x, nteams = TEAM ? cluster_count_x() : gridDim.x;
    double* my_scratch = scratch + (size_t)team * slot_scratch;
    double* my_z = zbuf + (size_t)blockIdx.x * slot_z;
    const unsigned int total = (unsigned)m.n_chains * (unsigned)m.n_lanes;
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
        // lanes in decreasing-work order, chains innermost: long lanes of every chain are scheduled first (LPT)
        const int lane_id = lane_order[t / m.n_chains];
        const int chain = t % m.n_chains;
        const unsigned gchain = (unsigned)(m.chain0 + chain);
        const int f = m.lane_factor[lane_id];
        double* theta = c.theta + (size_t)chain * m.n_params;
        const double* Ubase = c.Ueff + (size_t)chain * m.nU * m.n;
        for (int j = j0; j < j1; j++) {
            const uint32_t it = (uint32_t)(outer * m.nMH + j);
            for (int si = m.lane_off[lane_id]; si < m.lane_off[lane_id + 1]; si++) {
                const int s = m.lane_sites[si];
                const SiteDef sd = m.sites[s];
                // The scalar work of one update (proposal draw, three proposal/prior densities that depend on it, one that does not,
                // the acceptance uniform) is spread over the first lanes of three warps in two rounds instead of running as one
                // serial stretch on thread 0 while 255 threads wait: at n <= 256 it was ~9 % of the kernel.
                const double cur = theta[sd.param];
                const double shape_f = cur * cur / m.drift + 2.0, scale_f = cur * (shape_f - 1.0);
                if (threadIdx.x == 0) {
                    Stream st(m.seed, gchain, (uint32_t)s, stream_b(TAG_MH_PROP, it));
                    s_new = st.inv_gamma(shape_f, scale_f);
                } else if (threadIdx.x == 32) {
                    s_terms[0] = ig_logpdf(cur, sd.pshape, sd.pscale);
                } else if (threadIdx.x == 64) {
                    Stream sa(m.seed, gchain, (uint32_t)s, stream_b(TAG_MH_ACC, it));
                    s_terms[4] = log(sa.uniform());
                }
                __syncthreads();
                const double nw = s_new;
                if (threadIdx.x == 0) {
                    s_terms[1] = ig_logpdf(nw, shape_f, scale_f);                       // forward proposal density
                } else if (threadIdx.x == 32) {
                    const double shape_b = nw * nw / m.drift + 2.0, scale_b = nw * (shape_b - 1.0);
                    s_terms[2] = ig_logpdf(cur, shape_b, scale_b);                      // backward proposal density
                } else if (threadIdx.x == 64) {
                    s_terms[3] = ig_logpdf(nw, sd.pshape, sd.pscale);
                }
                double dlik;
                if (f >= 0) {
                    build_spec(m, c, chain, f, Ubase, sd.param, nw, &spec);
                    __syncthreads();
                    RbfGen gen{&spec, sm.exp2tab};
                    factor_run<RbfGen, TEAM>(gen, NCB, NCB, 1, my_scratch, my_z, sm, pipe, 1 << 30, nullptr, 0, nullptr, nullptr, 0, nullptr, /*keep_diag=*/false);
                    dlik = 0.0;
                } else {
                    __syncthreads();
                    dlik = 0.0;
                }
                if (threadIdx.x == 0) {
                    double lp_new = 0.0;
                    if (f >= 0) {
                        lp_new = logpdf_from(sm.out, m.n);
                        dlik = lp_new - c.lp[(size_t)chain * m.nF + f];
                    } else {
                        // uNoise: only the nU terms log N(U_k; 0, uNoise*SigmaU) change (closed form, cached q_k)
                        const double cur = theta[sd.param];
                        for (int k = 0; k < m.nU; k++)
                            dlik += -0.5 * (m.n * (log(nw) - log(cur)) + c.q[(size_t)chain * m.nU + k] * (1.0 / nw - 1.0 / cur));
                    }
                    const double dprior = s_terms[3] - s_terms[0];
                    const double s_part = dprior - s_terms[1] + s_terms[2];
                    const double alpha = s_part + dlik;
                    if (s_terms[4] < alpha && trank == 0) {
                        theta[sd.param] = nw;
                        if (f >= 0) c.lp[(size_t)chain * m.nF + f] = lp_new;
                        c.accepts[(size_t)chain * m.n_sites + s] += 1ull;
                    }
                }
                if constexpr (TEAM != 0) cluster_barrier();   // rank 0's update of theta / lp is visible to the team
                else __syncthreads();
            }
        }
    }
#ifdef GPSLC_PHASE_TIMING
    if (threadIdx.x == 0) atomicAdd(&g_phase_cycles[7], (unsigned long long)(clock64() - _k0));
#endif
}

#ifdef GPSLC_PHASE_TIMING
}  // namespace gpslc
extern "C" int gpslc_debug_phase_cycles(unsigned long long* out8, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out8, gpslc::g_phase_cycles, 64 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[64] = {0}; cudaMemcpyToSymbol(gpslc::g_phase_cycles, z, sizeof(z)); }
    return 0;
}
namespace gpslc {
#endif

// ------------------------------------------------------------------------------------------------ ESS over U_k
constexpr int ESS_MAX_EVALS = 200;
// ess[c][0..4] = log u, theta, theta_min, theta_max, next scalar-stream block
__device__ inline void ess_make_proposal(const ModelDev& m, const ChainDev& c, int chain, int k, double th, USmem& us) {
    const double* u = c.U + ((size_t)chain * m.nU + k) * m.n;
    const double* nu = c.nu + (size_t)chain * m.n;
    double* up = c.Uprop + (size_t)chain * m.n;
    const double cs = cos(th), sn = sin(th);
    for (int i = threadIdx.x; i < m.n; i += blockDim.x) up[i] = u[i] * cs + nu[i] * sn;
    __syncthreads();
    build_ueff(m, c.U + (size_t)chain * m.nU * m.n, k, up, c.UeffP + (size_t)chain * m.nU * m.n);
    const double qv = u_quad(m, c, chain, up, us);
    if (threadIdx.x == 0) { c.qP[chain] = qv; c.infoP[chain] = 0; }
}

__global__ void __launch_bounds__(256) ess_begin_kernel(ModelDev m, ChainDev c, int k, uint32_t it) {
    __shared__ USmem us;
    __shared__ double s_th;
    const int chain = blockIdx.x;
    const unsigned gchain = (unsigned)(m.chain0 + chain);
    const double un = c.theta[(size_t)chain * m.n_params + 0];
    Stream sn(m.seed, gchain, (uint32_t)k, stream_b(TAG_ESS_NU, it));
    u_prior_draw(m, c, chain, sn, un, c.nu + (size_t)chain * m.n);
    if (threadIdx.x == 0) {
        Stream ss(m.seed, gchain, (uint32_t)k, stream_b(TAG_ESS_SCALAR, it));
        double u, v;
        ss.uniform_pair(u, v);
        const double th = 2.0 * 3.14159265358979323846 * v;
        double* e = c.ess + (size_t)chain * 8;
        e[0] = log(u); e[1] = th; e[2] = th - 2.0 * 3.14159265358979323846; e[3] = th; e[4] = 1.0;
        s_th = th;
        c.active_a[chain] = chain;
        if (chain == 0) { c.n_active[0] = gridDim.x; c.n_active[1] = 0; }
    }
    __syncthreads();
    ess_make_proposal(m, c, chain, k, s_th, us);
}

__global__ void __launch_bounds__(256) ess_decide_kernel(ModelDev m, ChainDev c, int k, uint32_t it, const int* list_in,
                                                         int* list_out, unsigned int* n_out, const int* exist, int n_exist) {
    __shared__ USmem us;
    __shared__ double s_th;
    __shared__ int s_accept;
    const int chain = list_in[blockIdx.x];
    const unsigned gchain = (unsigned)(m.chain0 + chain);
    double* e = c.ess + (size_t)chain * 8;
    if (threadIdx.x == 0) {
        double w = 0.0;
        for (int i = 0; i < n_exist; i++) {
            const int f = exist[i];
            w += c.lpP[(size_t)chain * m.nF + f] - c.lp[(size_t)chain * m.nF + f];
        }
        // Gen's `elliptical_slice` compares the full `update` weight, which includes the N(0, uNoise*SigmaU) prior term
        // at the sliced address; ess_rule 1 = textbook likelihood-only rule.
        if (m.ess_rule == 0) w += -0.5 * (c.qP[chain] - c.q[(size_t)chain * m.nU + k]) / c.theta[(size_t)chain * m.n_params + 0];
        if (c.infoP[chain] != 0) w = -INFINITY;
        c.ess_evals[chain] += 1ull;
        // Gen loops `while weight <= log(u)`: a NaN weight leaves the loop; ESS_MAX_EVALS is a safety cap (oracle has the same)
        const bool acc = !(w <= e[0]) || (e[4] >= (double)ESS_MAX_EVALS);
        // capped on a non-PD proposal: keep the current state (s_accept = 2: stop slicing, copy nothing)
        s_accept = acc ? ((c.infoP[chain] != 0) ? 2 : 1) : 0;
        if (!acc) {
            double th = e[1];
            if (th < 0.0) e[2] = th; else e[3] = th;
            Stream ss(m.seed, gchain, (uint32_t)k, stream_b(TAG_ESS_SCALAR, it));
            ss.block = (uint32_t)e[4];
            th = e[2] + (e[3] - e[2]) * ss.uniform();
            e[4] += 1.0;
            e[1] = th;
            s_th = th;
            const unsigned int pos = atomicAdd(n_out, 1u);
            list_out[pos] = chain;
        }
    }
    __syncthreads();
    if (s_accept == 2) return;
    if (s_accept) {
        double* u = c.U + ((size_t)chain * m.nU + k) * m.n;
        const double* up = c.Uprop + (size_t)chain * m.n;
        for (int i = threadIdx.x; i < m.n; i += blockDim.x) u[i] = up[i];
        double* ue = c.Ueff + (size_t)chain * m.nU * m.n;
        const double* uep = c.UeffP + (size_t)chain * m.nU * m.n;
        for (int i = threadIdx.x; i < m.n * m.nU; i += blockDim.x) ue[i] = uep[i];
        for (int i = threadIdx.x; i < n_exist; i += blockDim.x) {
            const int f = exist[i];
            c.lp[(size_t)chain * m.nF + f] = c.lpP[(size_t)chain * m.nF + f];
        }
        if (threadIdx.x == 0) c.q[(size_t)chain * m.nU + k] = c.qP[chain];
    } else {
        ess_make_proposal(m, c, chain, k, s_th, us);
    }
}

// ------------------------------------------------------------------------------------------------ binary treatment
__device__ __forceinline__ double softplus(double x) { return fmax(x, 0.0) + log1p(exp(-fabs(x))); }
// sum_i log Bernoulli(T_i; expit(x_i)) (src/model_prior.jl:22-24) in the overflow-safe form
__device__ inline double block_bernoulli(const ModelDev& m, const double* Td, const double* f, const double* nu, double cs, double sn, double* red) {
    double part = 0.0;
    for (int i = threadIdx.x; i < m.n; i += blockDim.x) {
        const double x = f[i] * cs + nu[i] * sn;
        part += (Td[i] > 0.5) ? -softplus(-x) : -softplus(x);
    }
    return block_sum(part, red);
}

// mode 0: `generate` of logitT ~ N(0, K_T) = L z with the MODEL's T covariance (src/model_likelihood.jl:25-33,46-52,63-71)
// mode 1: once per outer iteration, factor the covariance of src/inference.jl:216-227 (per-dimension U vectors and the DATA X —
//         not the model's effective U / X, App. B1/B3) and draw the nES slice directions nu_j = L z_j used by the logitT
//         elliptical-slice updates of that iteration (App. B6: the covariance is NOT refreshed while U moves)
__global__ void __launch_bounds__(FTHREADS, CTAS_PER_SM)
logit_prior_kernel(ModelDev m, ChainDev c, int mode, int outer, double* scratch, size_t slot_scratch, double* zbuf, size_t slot_z,
                   double* xibuf, unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    __shared__ RbfSpec spec;
    __shared__ unsigned int job;
    factor_smem_init(sm);
    Pipe pipe{0, 0};
    const int NCB = ceil_div(m.n, NB), FT = m.nX;
    double* my_scratch = scratch + (size_t)blockIdx.x * slot_scratch;
    double* my_z = zbuf + (size_t)blockIdx.x * slot_z;
    double* xi = xibuf + (size_t)blockIdx.x * 4 * m.n;
    for (;;) {
        if (threadIdx.x == 0) job = atomicAdd(counter, 1u);
        __syncthreads();
        const int chain = (int)job;
        if (chain >= m.n_chains) break;
        const unsigned gchain = (unsigned)(m.chain0 + chain);
        build_spec(m, c, chain, FT, c.Ueff + (size_t)chain * m.nU * m.n, -1, 0.0, &spec);
        __syncthreads();
        if (mode == 1) {
            // override the feature columns: U_k vectors as stored (column-wise) and the observed X
            const FactorDef& fd = m.fdef[FT];
            for (int d = threadIdx.x; d < fd.D; d += blockDim.x) {
                const int kind = fd.src_kind[d], idx = fd.src_idx[d];
                spec.feat[d] = (kind == SRC_U) ? c.U + ((size_t)chain * m.nU + idx) * m.n : m.X + (size_t)chain * m.xstride + (size_t)idx * m.n;
            }
            __syncthreads();
        }
        RbfGen gen{&spec, sm.exp2tab};
        factor_run(gen, NCB, NCB, 0, my_scratch, my_z, sm, pipe);
        if (threadIdx.x == 0 && sm.out.info != 0) atomicMax(&c.info[chain], sm.out.info);
        const int ndraw = (mode == 0) ? 1 : m.nES;
        for (int s0 = 0; s0 < ndraw; s0 += 4) {
            const int ns = min(4, ndraw - s0);
            __syncthreads();
            for (int e = threadIdx.x; e < ns * m.n; e += blockDim.x) {
                const int s = e / m.n, col = e - s * m.n;
                Stream st(m.seed, gchain, (uint32_t)m.nU,
                          mode == 0 ? stream_b(TAG_INIT_VEC, 0) : stream_b(TAG_ESS_NU, (uint32_t)(outer * m.nES + s0 + s)));
                xi[(size_t)s * m.n + col] = st.normal_at(col);
            }
            __syncthreads();
            for (int i = threadIdx.x; i < m.n; i += blockDim.x) {
                double accv[4];
                tri_matvec_row<4>(my_scratch, NCB, 0, 0, i, m.n, xi, m.n, ns, accv);
                for (int s = 0; s < ns; s++) {
                    if (mode == 0) c.logitT[(size_t)chain * m.n + i] = accv[s];
                    else c.nuL[((size_t)chain * m.nES + s0 + s) * m.n + i] = accv[s];
                }
            }
        }
        __syncthreads();
    }
}

// `elliptical_slice(trace, :logitT, zeros(n), logitTCov)` (src/inference.jl:233,293,347) for every chain: K_T is fixed during
// the slice, so ONE factorisation with right-hand sides (f, nu) gives the quadratic form on the whole ellipse,
// f'K^-1 f cos^2 + 2 f'K^-1 nu sin cos + nu'K^-1 nu sin^2; each shrink step then costs only the O(n) Bernoulli terms.
__global__ void __launch_bounds__(FTHREADS, CTAS_PER_SM)
ess_logit_kernel(ModelDev m, ChainDev c, int jj, uint32_t it, double* scratch, size_t slot_scratch, double* zbuf, size_t slot_z,
                 unsigned int* counter) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    __shared__ RbfSpec spec;
    __shared__ unsigned int job;
    __shared__ double red[32];
    __shared__ double s_cs, s_sn;
    __shared__ int s_done;
    factor_smem_init(sm);
    Pipe pipe{0, 0};
    const int NCB = ceil_div(m.n, NB), FT = m.nX;
    double* my_scratch = scratch + (size_t)blockIdx.x * slot_scratch;
    double* my_z = zbuf + (size_t)blockIdx.x * slot_z;
    for (;;) {
        if (threadIdx.x == 0) job = atomicAdd(counter, 1u);
        __syncthreads();
        const int chain = (int)job;
        if (chain >= m.n_chains) break;
        const unsigned gchain = (unsigned)(m.chain0 + chain);
        double* f = c.logitT + (size_t)chain * m.n;
        const double* nu = c.nuL + ((size_t)chain * m.nES + jj) * m.n;
        build_spec(m, c, chain, FT, c.Ueff + (size_t)chain * m.nU * m.n, -1, 0.0, &spec);
        __syncthreads();
        if (threadIdx.x == 0) spec.y[1] = nu;
        __syncthreads();
        RbfGen gen{&spec, sm.exp2tab};
        factor_run(gen, NCB, NCB, 2, my_scratch, my_z, sm, pipe);
        const FactorOut o = sm.out;
        const double* Td = m.T + (size_t)chain * m.tstride;
        const double bern0 = block_bernoulli(m, Td, f, nu, 1.0, 0.0, red);
        double logu = 0.0, th = 0.0, tmin = 0.0, tmax = 0.0;
        Stream ss(m.seed, gchain, (uint32_t)m.nU, stream_b(TAG_ESS_SCALAR, it));
        if (threadIdx.x == 0) {
            double u, v;
            ss.uniform_pair(u, v);
            logu = log(u); th = 2.0 * 3.14159265358979323846 * v; tmin = th - 2.0 * 3.14159265358979323846; tmax = th;
            s_cs = cos(th); s_sn = sin(th); s_done = (o.info != 0) ? 2 : 0;
        }
        __syncthreads();
        int evals = 0;
        double quad = o.gram[0];
        while (s_done == 0) {
            const double cs = s_cs, sn = s_sn;
            const double bern = block_bernoulli(m, Td, f, nu, cs, sn, red);
            __syncthreads();
            if (threadIdx.x == 0) {
                evals++;
                quad = cs * cs * o.gram[0] + 2.0 * sn * cs * o.gram[1] + sn * sn * o.gram[2];
                // ess_rule 0: Gen's full `update` weight = change of the N(0, K_T) term at :logitT + change of the Bernoulli terms;
                // ess_rule 1: textbook slice test on the likelihood (the Bernoulli terms) only — same switch as for the U_k
                const double w = (m.ess_rule == 0 ? -0.5 * (quad - o.gram[0]) : 0.0) + (bern - bern0);
                if (!(w <= logu) || evals >= ESS_MAX_EVALS) {
                    s_done = 1;
                } else {
                    if (th < 0.0) tmin = th; else tmax = th;
                    th = tmin + (tmax - tmin) * ss.uniform();
                    s_cs = cos(th); s_sn = sin(th);
                }
            }
            __syncthreads();
        }
        if (s_done == 1) {
            const double cs = s_cs, sn = s_sn;
            for (int i = threadIdx.x; i < m.n; i += blockDim.x) f[i] = f[i] * cs + nu[i] * sn;
            if (threadIdx.x == 0) {
                c.lp[(size_t)chain * m.nF + FT] = -0.5 * (m.n * LOG_2PI + o.logdet + quad);
                c.ess_evals_logit[chain] += (unsigned long long)evals;
            }
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------ sample record
__global__ void record_kernel(ModelDev m, ChainDev c, double* samples, int slot) {
    const int chain = blockIdx.x;
    double* dst = samples + ((size_t)slot * m.n_chains + chain) * m.stride;
    const double* theta = c.theta + (size_t)chain * m.n_params;
    for (int i = threadIdx.x; i < m.n_params; i += blockDim.x) dst[i] = theta[i];
    dst += m.n_params;
    const int nu = m.nU * m.n;
    const double* U = c.U + (size_t)chain * nu;
    for (int i = threadIdx.x; i < nu; i += blockDim.x) dst[i] = U[i];
    dst += nu;
    if (m.binary) {
        const double* lt = c.logitT + (size_t)chain * m.n;
        for (int i = threadIdx.x; i < m.n; i += blockDim.x) dst[i] = lt[i];
        dst += m.n;
    }
    if (m.has_xmodel) {
        const double* xm = c.Xmodel + (size_t)chain * m.n * m.nX;
        for (int i = threadIdx.x; i < m.n * m.nX; i += blockDim.x) dst[i] = xm[i];
    }
}

// ================================================================================================ host side

template <class T>
static int dev_alloc(Sampler* s, T** p, size_t count) {
    *p = nullptr;
    if (count == 0) count = 1;
    GP_CUDA(s->ctx, s->ctx->block_alloc((void**)p, count * sizeof(T)));
    s->owned.push_back({(void*)*p, count * sizeof(T)});
    return GPSLC_OK;
}
template <class T>
static int dev_upload(Sampler* s, T** p, const std::vector<T>& v) {
    GP_TRY(dev_alloc(s, p, v.size()));
    if (!v.empty()) GP_CUDA(s->ctx, cudaMemcpyAsync(*p, v.data(), v.size() * sizeof(T), cudaMemcpyHostToDevice, s->ctx->stream));
    return GPSLC_OK;
}

static int param_idx(int nX, int nU, const char* name, int i, int j) {
    std::string s(name);
    if (s == "uNoise") return 0;
    if (s == "tNoise") return 1;
    if (s == "yNoise") return 2;
    if (s == "tyLS") return 3;
    if (s == "tScale") return 4;
    if (s == "yScale") return 5;
    if (s == "xNoise") return 6 + i;
    if (s == "xScale") return 6 + nX + i;
    if (s == "xtLS") return 6 + 2 * nX + i;
    if (s == "xyLS") return 6 + 3 * nX + i;
    if (s == "utLS") return 6 + 4 * nX + i;
    if (s == "uyLS") return 6 + 4 * nX + nU + i;
    return 6 + 4 * nX + 2 * nU + i * nX + j;  // uxLS
}

void sampler_free(Sampler* s) {
    if (!s) return;
    cudaStreamSynchronize(s->ctx->stream);
    for (auto& b : s->owned) s->ctx->block_free(b.first, b.second);
    if (s->samples) s->ctx->block_free(s->samples, s->samples_bytes);
    delete s;
}

// Build every table. data pointers are host pointers (loc==0) or device pointers (loc==1).
int sampler_create(Ctx* ctx, int loc, int n, int nX, int nU, int binary, const double* X, const double* T, const double* Y,
                   int n_obj, const int* obj_counts, const double* sigma_dense, double eps, double cov, const double* pshape, const double* pscale,
                   double drift, int nMH, int nES, int n_chains, unsigned long long seed, int chain_offset, int u_layout_mode,
                   int ess_rule, int observe_x, int per_chain_data, Sampler** out) {
    *out = nullptr;
    if (n <= 0 || nX < 0 || nU < 0 || n_chains <= 0 || !T || !Y || (nX > 0 && !X)) return ctx->fail(GPSLC_ERR_ARG, "sampler_create: bad argument");
    if (nU + nX + 1 > DMAX) return ctx->fail(GPSLC_ERR_UNSUPPORTED, "sampler_create: nU + nX + 1 exceeds DMAX");
    const bool has_u = nU > 0, has_x = nX > 0;
    // InvGamma(shape, scale) priors and the proposal variance must be positive and finite (Distributions.InverseGamma throws
    // an ArgumentError otherwise); a bad value would otherwise reach the gamma sampler inside a kernel
    for (int fam = 0; fam < P_NFAM; fam++)
        if (!(pshape[fam] > 0.0) || !(pscale[fam] > 0.0) || !(pshape[fam] < 1e300) || !(pscale[fam] < 1e300))
            return ctx->fail(GPSLC_ERR_ARG, "sampler_create: InvGamma prior shape and scale must be positive and finite (family " + std::to_string(fam) + ")");
    if (!(drift > 0.0) || !(drift < 1e300)) return ctx->fail(GPSLC_ERR_ARG, "sampler_create: drift (proposal variance) must be positive and finite");
    const bool dense_u = has_u && sigma_dense != nullptr && (n_obj <= 0 || !obj_counts);
    if (has_u && !dense_u) {
        if (n_obj <= 0 || !obj_counts) return ctx->fail(GPSLC_ERR_ARG, "sampler_create: nU > 0 needs SigmaU: its object counts, or the dense matrix");
        long long tot = 0;
        for (int o = 0; o < n_obj; o++) { if (obj_counts[o] <= 0) return ctx->fail(GPSLC_ERR_ARG, "sampler_create: non-positive object count"); tot += obj_counts[o]; }
        if (tot != n) return ctx->fail(GPSLC_ERR_ARG, "sampler_create: object counts do not sum to n");
        if (!((1.0 + eps) - cov > 0.0) || !(cov >= 0.0)) return ctx->fail(GPSLC_ERR_ARG, "sampler_create: SigmaU needs 0 <= cov < 1+eps");
    }
    GP_CUDA(ctx, cudaSetDevice(ctx->device));
    Sampler* s = new Sampler();
    s->ctx = ctx;
    ModelDev& m = s->m;
    m.n = n; m.nU = nU; m.nX = nX; m.nF = nX + 2; m.binary = binary;
    m.n_params = 6 + 4 * nX + 2 * nU + nU * nX;
    m.has_xmodel = (!has_u && has_x && !observe_x) ? 1 : 0;
    m.stride = m.n_params + nU * n + (binary ? n : 0) + (m.has_xmodel ? n * nX : 0);
    m.u_layout_reference = (u_layout_mode == 0) ? 1 : 0;
    m.ess_rule = ess_rule;
    m.eps = eps; m.cov = cov; m.dU = (1.0 + eps) - cov; m.drift = drift;
    { const char* e = getenv("GPSLC_LS_UNSQUARED"); m.ls_unsquared = (e && atoi(e) == 1) ? 1 : 0; }
    m.seed = seed; m.chain0 = chain_offset; m.n_chains = n_chains;
    m.nMH = (!has_u && !has_x) ? 1 : nMH;   // inference.jl:157-160: three sites once per outer iteration
    m.nES = nES;
    m.n_obj = (has_u && !dense_u) ? n_obj : 0;
    m.dense_u = dense_u ? 1 : 0;
    m.SL = nullptr;

    // ---- factor table (src/model_likelihood.jl)
    s->h_fdef.assign(m.nF, FactorDef{});
    auto uxls_param = [&](int k, int c) {  // lengthscale of U column c in the X_k kernel (model_prior.jl:110, App. B1)
        if (m.u_layout_reference) { const long long f = (long long)k + (long long)c * nX; return param_idx(nX, nU, "uxLS", (int)(f % nU), (int)(f / nU)); }
        return param_idx(nX, nU, "uxLS", c, k);
    };
    for (int f = 0; f < m.nF; f++) {
        FactorDef& fd = s->h_fdef[f];
        int D = 0;
        if (f < nX) {
            fd.exists = has_u;
            for (int c = 0; c < nU; c++) { fd.src_kind[D] = SRC_U; fd.src_idx[D] = c; fd.ls_param[D] = uxls_param(f, c); D++; }
            fd.scale_param = param_idx(nX, nU, "xScale", f, 0); fd.noise_param = param_idx(nX, nU, "xNoise", f, 0);
            fd.target_kind = TGT_XCOL; fd.target_idx = f;
        } else if (f == nX) {
            fd.exists = has_u || has_x;
            for (int c = 0; c < nU; c++) { fd.src_kind[D] = SRC_U; fd.src_idx[D] = c; fd.ls_param[D] = param_idx(nX, nU, "utLS", c, 0); D++; }
            for (int k = 0; k < nX; k++) { fd.src_kind[D] = SRC_X; fd.src_idx[D] = k; fd.ls_param[D] = param_idx(nX, nU, "xtLS", k, 0); D++; }
            fd.scale_param = 4; fd.noise_param = 1;
            fd.target_kind = binary ? TGT_LOGIT : TGT_T; fd.target_idx = 0;
        } else {
            fd.exists = 1;
            for (int c = 0; c < nU; c++) { fd.src_kind[D] = SRC_U; fd.src_idx[D] = c; fd.ls_param[D] = param_idx(nX, nU, "uyLS", c, 0); D++; }
            for (int k = 0; k < nX; k++) { fd.src_kind[D] = SRC_X; fd.src_idx[D] = k; fd.ls_param[D] = param_idx(nX, nU, "xyLS", k, 0); D++; }
            fd.src_kind[D] = SRC_T; fd.src_idx[D] = 0; fd.ls_param[D] = 3; D++;
            fd.scale_param = 5; fd.noise_param = 2;
            fd.target_kind = TGT_Y; fd.target_idx = 0;
        }
        fd.D = D;
    }
    // ---- MH sites in sweep order (src/inference.jl:23-44, 76-87, 127-137, 158-160)
    auto add_site = [&](const char* name, int i, int j, int fam, int factor) {
        SiteDef sd; sd.param = param_idx(nX, nU, name, i, j); sd.factor = factor; sd.pshape = pshape[fam]; sd.pscale = pscale[fam];
        s->h_sites.push_back(sd);
    };
    const int FT = nX, FY = nX + 1;
    if (has_u) add_site("uNoise", 0, 0, P_UNOISE, -1);
    if (has_u || has_x) add_site("tNoise", 0, 0, P_TNOISE, FT);
    add_site("yNoise", 0, 0, P_YNOISE, FY);
    add_site("tyLS", 0, 0, P_TYLS, FY);
    if (has_u) for (int k = 0; k < nU; k++) {
        add_site("utLS", k, 0, P_UTLS, FT);
        add_site("uyLS", k, 0, P_UYLS, FY);
        for (int l = 0; l < nX; l++) {
            const int fx = m.u_layout_reference ? (int)(((long long)k + (long long)l * nU) % nX) : l;
            add_site("uxLS", k, l, P_UXLS, fx);
        }
    }
    if (has_x) for (int k = 0; k < nX; k++) {
        if (has_u) add_site("xNoise", k, 0, P_XNOISE, k);
        add_site("xtLS", k, 0, P_XTLS, FT);
        add_site("xyLS", k, 0, P_XYLS, FY);
        if (has_u) add_site("xScale", k, 0, P_XSCALE, k);
    }
    if (has_u || has_x) add_site("tScale", 0, 0, P_TSCALE, FT);
    add_site("yScale", 0, 0, P_YSCALE, FY);
    m.n_sites = (int)s->h_sites.size();
    // ---- lanes: sites grouped by the factor they feed (-1 = uNoise lane)
    {
        std::vector<int> keys;
        for (auto& sd : s->h_sites) { bool seen = false; for (int k : keys) seen |= (k == sd.factor); if (!seen) keys.push_back(sd.factor); }
        s->h_lane_off.push_back(0);
        for (int key : keys) {
            for (int i = 0; i < m.n_sites; i++) if (s->h_sites[i].factor == key) s->h_lane_sites.push_back(i);
            s->h_lane_off.push_back((int)s->h_lane_sites.size());
            s->h_lane_factor.push_back(key);
        }
        m.n_lanes = (int)keys.size();
        s->lane_task_order.resize(m.n_lanes);
        for (int l = 0; l < m.n_lanes; l++) s->lane_task_order[l] = l;
        auto work = [&](int l) { return (s->h_lane_factor[l] < 0 ? 0 : 1000) + (s->h_lane_off[l + 1] - s->h_lane_off[l]); };
        std::sort(s->lane_task_order.begin(), s->lane_task_order.end(), [&](int a, int b) { return work(a) > work(b); });
    }
    std::vector<int> exist;
    for (int f = 0; f < m.nF; f++) if (s->h_fdef[f].exists) exist.push_back(f);
    s->n_exist = (int)exist.size();

    // ---- uploads
    int rc = GPSLC_OK;
    auto up_data = [&](const double* src, size_t count, const double** dst) -> int {
        if (!src || count == 0) { *dst = nullptr; return GPSLC_OK; }
        if (loc == 1) { *dst = src; return GPSLC_OK; }
        double* d;
        GP_TRY(dev_alloc(s, &d, count));
        GP_CUDA(ctx, cudaMemcpyAsync(d, src, count * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
        *dst = d;
        return GPSLC_OK;
    };
#define S_TRY(expr) do { rc = (expr); if (rc) { sampler_free(s); return rc; } } while (0)
    const size_t ncopies = per_chain_data ? (size_t)n_chains : 1;
    m.xstride = per_chain_data ? (size_t)n * nX : 0;
    m.tstride = per_chain_data ? (size_t)n : 0;
    S_TRY(up_data(X, ncopies * n * nX, &m.X));
    S_TRY(up_data(T, ncopies * n, &m.T));
    S_TRY(up_data(Y, ncopies * n, &m.Y));
    {
        std::vector<int> start(m.n_obj + 1, 0), of(n, 0);
        for (int o = 0; o < m.n_obj; o++) { start[o + 1] = start[o] + obj_counts[o]; for (int i = start[o]; i < start[o + 1]; i++) of[i] = o; }
        int *d1, *d2;
        S_TRY(dev_upload(s, &d1, start)); S_TRY(dev_upload(s, &d2, of));
        m.obj_start = d1; m.obj_of = d2;
    }
    { FactorDef* d; S_TRY(dev_upload(s, &d, s->h_fdef)); m.fdef = d; }
    { SiteDef* d; S_TRY(dev_upload(s, &d, s->h_sites)); m.sites = d; }
    { int* d; S_TRY(dev_upload(s, &d, s->h_lane_sites)); m.lane_sites = d; }
    { int* d; S_TRY(dev_upload(s, &d, s->h_lane_off)); m.lane_off = d; }
    { int* d; S_TRY(dev_upload(s, &d, s->h_lane_factor)); m.lane_factor = d; }
    S_TRY(dev_upload(s, &s->d_lane_order, s->lane_task_order));
    S_TRY(dev_upload(s, &s->d_exist, exist));
    // ---- chain state
    ChainDev& c = s->c;
    const size_t C = n_chains;
    S_TRY(dev_alloc(s, &c.theta, C * m.n_params));
    S_TRY(dev_alloc(s, &c.U, C * nU * n));
    S_TRY(dev_alloc(s, &c.Ueff, C * nU * n));
    S_TRY(dev_alloc(s, &c.UeffP, C * nU * n));
    S_TRY(dev_alloc(s, &c.Uprop, C * n));
    S_TRY(dev_alloc(s, &c.nu, C * n));
    S_TRY(dev_alloc(s, &c.lp, C * m.nF));
    S_TRY(dev_alloc(s, &c.lpP, C * m.nF));
    S_TRY(dev_alloc(s, &c.q, C * (nU > 0 ? nU : 1)));
    S_TRY(dev_alloc(s, &c.qP, C));
    S_TRY(dev_alloc(s, &c.ess, C * 8));
    S_TRY(dev_alloc(s, &c.logitT, binary ? C * n : 1));
    S_TRY(dev_alloc(s, &c.Xmodel, m.has_xmodel ? C * n * nX : 1));
    S_TRY(dev_alloc(s, &c.info, C));
    S_TRY(dev_alloc(s, &c.infoP, C));
    S_TRY(dev_alloc(s, &c.active_a, C));
    S_TRY(dev_alloc(s, &c.active_b, C));
    S_TRY(dev_alloc(s, &c.n_active, 2));
    S_TRY(dev_alloc(s, &c.accepts, C * m.n_sites));
    S_TRY(dev_alloc(s, &c.ess_evals, C));
    S_TRY(dev_alloc(s, &c.ess_evals_logit, C));
    S_TRY(dev_alloc(s, &c.nuL, binary ? C * (size_t)(nES > 0 ? nES : 1) * n : 1));
    S_TRY(dev_alloc(s, &c.zs, dense_u ? C * (size_t)ceil_div(n, NB) * NB : 1));
    cudaMemsetAsync(c.lp, 0, C * m.nF * sizeof(double), ctx->stream);
    cudaMemsetAsync(c.lpP, 0, C * m.nF * sizeof(double), ctx->stream);
    cudaMemsetAsync(c.accepts, 0, C * m.n_sites * sizeof(unsigned long long), ctx->stream);
    cudaMemsetAsync(c.ess_evals, 0, C * sizeof(unsigned long long), ctx->stream);
    cudaMemsetAsync(c.ess_evals_logit, 0, C * sizeof(unsigned long long), ctx->stream);
    cudaMemsetAsync(c.info, 0, C * sizeof(int), ctx->stream);
    cudaMemsetAsync(c.infoP, 0, C * sizeof(int), ctx->stream);
    const int NCB = ceil_div(n, NB);
    { int g0 = 0; S_TRY(ensure_workspace(ctx, NCB, NCB, (long long)n_chains * (nX + 2), &g0)); }
    if (dense_u) {
        // SigmaU -> device (host pointer unless loc == 1), factor once; a non-PD SigmaU is the PosDefException generateU would throw
        const double* dS = sigma_dense;
        if (loc != 1) { double* t; S_TRY(dev_alloc(s, &t, (size_t)n * n)); cudaMemcpyAsync(t, sigma_dense, (size_t)n * n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream); dS = t; }
        double* SL; int* dinfo;
        S_TRY(dev_alloc(s, &SL, scratch_doubles(NCB, NCB)));
        S_TRY(dev_alloc(s, &dinfo, 1));
        cudaFuncSetAttribute(sigma_u_factor_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem));
        sigma_u_factor_kernel<<<1, FTHREADS, sizeof(FactorSmem), ctx->stream>>>(dS, n, SL, ctx->zbuf, dinfo);
        ctx->launches++;
        int hinfo = 0;
        cudaError_t e1 = cudaMemcpyAsync(&hinfo, dinfo, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream);
        cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
        if (e1 != cudaSuccess || e2 != cudaSuccess) { sampler_free(s); return ctx->cuda_fail(e1 != cudaSuccess ? e1 : e2, "sigma_u_factor_kernel"); }
        if (hinfo != 0) { sampler_free(s); return ctx->fail(GPSLC_ERR_NOT_PD, "SigmaU is not positive definite (leading minor " + std::to_string(hinfo) + ")"); }
        m.SL = SL;
    }
#undef S_TRY
    *out = s;
    return GPSLC_OK;
}

// Launch one of the two instantiations of a factor kernel: one CTA per task, or (few large tasks) one thread-block cluster of
// `team` CTAs per task. *grid is the number of workspace slots (= teams); it is clipped to what can be resident.
template <class K1, class KT, class... Args>
static int launch_single_or_team(Ctx* ctx, K1 k_single, KT k_team, int team, int grid, const char* what, Args... args) {
    if (team == 1) {
        GP_CUDA(ctx, cudaFuncSetAttribute(k_single, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
        GP_CUDA(ctx, cudaFuncSetAttribute(k_single, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        k_single<<<grid, FTHREADS, sizeof(FactorSmem), ctx->stream>>>(args...);
        ctx->launches++;
        GP_CUDA(ctx, cudaGetLastError());
        return GPSLC_OK;
    }
    GP_CUDA(ctx, cudaFuncSetAttribute(k_team, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = team; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(FTHREADS); cfg.dynamicSmemBytes = sizeof(FactorSmem); cfg.stream = ctx->stream;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cfg.gridDim = dim3(grid * team);
    int max_clusters = 0;
    GP_CUDA(ctx, cudaOccupancyMaxActiveClusters(&max_clusters, k_team, &cfg));
    if (max_clusters < 1) return ctx->fail(GPSLC_ERR_CUDA, std::string(what) + ": no resident cluster of the requested size");
    if (grid > max_clusters) grid = max_clusters;   // fewer teams than slots: the teams loop over the tasks
    cfg.gridDim = dim3(grid * team);
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_team, args...);
    ctx->launches++;
    if (e != cudaSuccess) return ctx->cuda_fail(e, what);
    return GPSLC_OK;
}

static int launch_eval(Sampler* s, const int* list, const unsigned int* n_list_dev, int n_list_host, int proposed) {
    Ctx* ctx = s->ctx;
    if (s->n_exist == 0 || n_list_host == 0) return GPSLC_OK;
    GP_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
    const long long tasks = (long long)n_list_host * s->n_exist;
    const int NCB = ceil_div(s->m.n, NB);
    const int team = pick_team(ctx, tasks, NCB);
    int grid = 0;
    GP_TRY(ensure_workspace(ctx, NCB, NCB, tasks, &grid, team));
    return launch_single_or_team(ctx, eval_factors_kernel<0>, eval_factors_kernel<1>, team, grid, "eval_factors_kernel", s->m, s->c, list,
                                 n_list_dev, s->m.n_chains, (const int*)s->d_exist, s->n_exist, proposed, ctx->scratch, ctx->slot_scratch_d,
                                 ctx->zbuf, ctx->slot_z_d, ctx->counter);
}

static int launch_logit_prior(Sampler* s, int mode, int outer) {
    Ctx* ctx = s->ctx;
    GP_CUDA(ctx, cudaFuncSetAttribute(logit_prior_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
    GP_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
    int grid = 0;
    GP_TRY(ensure_workspace(ctx, ceil_div(s->m.n, NB), ceil_div(s->m.n, NB), s->m.n_chains, &grid));
    if (!s->xibuf) {
        const size_t xb = (size_t)ctx->slots * 4 * s->m.n * sizeof(double);
        GP_CUDA(ctx, ctx->block_alloc((void**)&s->xibuf, xb));
        s->owned.push_back({(void*)s->xibuf, xb});
    }
    logit_prior_kernel<<<grid, FTHREADS, sizeof(FactorSmem), ctx->stream>>>(s->m, s->c, mode, outer, ctx->scratch, ctx->slot_scratch_d,
                                                                          ctx->zbuf, ctx->slot_z_d, s->xibuf, ctx->counter);
    ctx->launches++;
    GP_CUDA(ctx, cudaGetLastError());
    return GPSLC_OK;
}

static int launch_ess_logit(Sampler* s, int jj, uint32_t it) {
    Ctx* ctx = s->ctx;
    GP_CUDA(ctx, cudaFuncSetAttribute(ess_logit_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FactorSmem)));
    GP_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
    int grid = 0;
    GP_TRY(ensure_workspace(ctx, ceil_div(s->m.n, NB), ceil_div(s->m.n, NB), s->m.n_chains, &grid));
    ess_logit_kernel<<<grid, FTHREADS, sizeof(FactorSmem), ctx->stream>>>(s->m, s->c, jj, it, ctx->scratch, ctx->slot_scratch_d, ctx->zbuf,
                                                                        ctx->slot_z_d, ctx->counter);
    ctx->launches++;
    GP_CUDA(ctx, cudaGetLastError());
    return GPSLC_OK;
}

// `generate`: prior draws + initial factor log-densities. A non-PD initial factor is an error like the reference's
// PosDefException.
int sampler_init(Sampler* s) {
    Ctx* ctx = s->ctx;
    init_chains_kernel<<<s->m.n_chains, 256, 0, ctx->stream>>>(s->m, s->c);
    ctx->launches++;
    GP_CUDA(ctx, cudaGetLastError());
    if (s->m.binary && s->h_fdef[s->m.nX].exists) GP_TRY(launch_logit_prior(s, 0, 0));
    GP_TRY(launch_eval(s, nullptr, nullptr, s->m.n_chains, 0));
    std::vector<int> info(s->m.n_chains);
    GP_CUDA(ctx, cudaMemcpyAsync(info.data(), s->c.info, info.size() * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (size_t i = 0; i < info.size(); i++)
        if (info[i] != 0) return ctx->fail(GPSLC_ERR_NOT_PD, "initial state: covariance not positive definite (chain " + std::to_string(i) + ", minor " + std::to_string(info[i]) + ")");
    s->outer_done = 0; s->sweeps_done = 0;
    return GPSLC_OK;
}

// overwrite chain state from packed records [C][stride] (theta | U | ...) and recompute caches
int sampler_set_state(Sampler* s, int loc, const double* packed) {
    Ctx* ctx = s->ctx;
    const ModelDev& m = s->m;
    const cudaMemcpyKind kind = loc == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    GP_CUDA(ctx, cudaMemcpy2DAsync(s->c.theta, m.n_params * sizeof(double), packed, m.stride * sizeof(double), m.n_params * sizeof(double), m.n_chains, kind, ctx->stream));
    if (m.nU > 0)
        GP_CUDA(ctx, cudaMemcpy2DAsync(s->c.U, (size_t)m.nU * m.n * sizeof(double), packed + m.n_params, m.stride * sizeof(double), (size_t)m.nU * m.n * sizeof(double), m.n_chains, kind, ctx->stream));
    if (m.binary)
        GP_CUDA(ctx, cudaMemcpy2DAsync(s->c.logitT, (size_t)m.n * sizeof(double), packed + m.n_params + m.nU * m.n, m.stride * sizeof(double), (size_t)m.n * sizeof(double), m.n_chains, kind, ctx->stream));
    if (m.has_xmodel)
        GP_CUDA(ctx, cudaMemcpy2DAsync(s->c.Xmodel, (size_t)m.nX * m.n * sizeof(double), packed + m.n_params + m.nU * m.n + (m.binary ? m.n : 0), m.stride * sizeof(double), (size_t)m.nX * m.n * sizeof(double), m.n_chains, kind, ctx->stream));
    refresh_chains_kernel<<<m.n_chains, 256, 0, ctx->stream>>>(s->m, s->c);
    ctx->launches++;
    GP_CUDA(ctx, cudaGetLastError());
    GP_TRY(launch_eval(s, nullptr, nullptr, m.n_chains, 0));
    GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GPSLC_OK;
}

int sampler_mh(Sampler* s, int outer, int j0, int j1) {
    Ctx* ctx = s->ctx;
    if (j1 <= j0) return GPSLC_OK;
    GP_CUDA(ctx, cudaMemsetAsync(ctx->counter, 0, sizeof(unsigned int), ctx->stream));
    const long long tasks = (long long)s->m.n_chains * s->m.n_lanes;
    const int NCB = ceil_div(s->m.n, NB);
    const int team = pick_team(ctx, tasks, NCB);
    int grid = 0;
    GP_TRY(ensure_workspace(ctx, NCB, NCB, tasks, &grid, team));
    return launch_single_or_team(ctx, mh_lanes_kernel<0>, mh_lanes_kernel<1>, team, grid, "mh_lanes_kernel", s->m, s->c,
                                 (const int*)s->d_lane_order, outer, j0, j1, ctx->scratch, ctx->slot_scratch_d, ctx->zbuf, ctx->slot_z_d,
                                 ctx->counter);
}

// one elliptical-slice update of U_k for every chain (src/inference.jl:50-54): host loop over shrink rounds
int sampler_ess_u(Sampler* s, int k, uint32_t it) {
    Ctx* ctx = s->ctx;
    const ModelDev& m = s->m;
    ess_begin_kernel<<<m.n_chains, 256, 0, ctx->stream>>>(s->m, s->c, k, it);
    ctx->launches++;
    GP_CUDA(ctx, cudaGetLastError());
    int cur = 0;
    unsigned int n_act = (unsigned)m.n_chains;
    while (n_act > 0) {
        int* list_in = cur == 0 ? s->c.active_a : s->c.active_b;
        int* list_out = cur == 0 ? s->c.active_b : s->c.active_a;
        GP_TRY(launch_eval(s, list_in, s->c.n_active + cur, (int)n_act, 1));
        GP_CUDA(ctx, cudaMemsetAsync(s->c.n_active + (1 - cur), 0, sizeof(unsigned int), ctx->stream));
        ess_decide_kernel<<<n_act, 256, 0, ctx->stream>>>(s->m, s->c, k, it, list_in, list_out, s->c.n_active + (1 - cur), s->d_exist, s->n_exist);
        ctx->launches++;
        GP_CUDA(ctx, cudaGetLastError());
        cur = 1 - cur;
        GP_CUDA(ctx, cudaMemcpyAsync(&n_act, s->c.n_active + cur, sizeof(unsigned int), cudaMemcpyDeviceToHost, ctx->stream));
        GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    }
    return GPSLC_OK;
}

int sampler_reserve_samples(Sampler* s, int n_outer_total) {
    Ctx* ctx = s->ctx;
    if (n_outer_total <= s->samples_cap) return GPSLC_OK;
    double* nw = nullptr;
    const size_t per = (size_t)s->m.n_chains * s->m.stride;
    const size_t nbytes = per * n_outer_total * sizeof(double);
    GP_CUDA(ctx, ctx->block_alloc((void**)&nw, nbytes));
    if (s->samples) {
        GP_CUDA(ctx, cudaMemcpyAsync(nw, s->samples, per * s->outer_done * sizeof(double), cudaMemcpyDeviceToDevice, ctx->stream));
        GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->block_free(s->samples, s->samples_bytes);
    }
    s->samples = nw;
    s->samples_bytes = nbytes;
    s->samples_cap = n_outer_total;
    return GPSLC_OK;
}

// run `n_outer` more outer iterations (MH sweeps + ESS passes + record)
int sampler_run(Sampler* s, int n_outer) {
    Ctx* ctx = s->ctx;
    const ModelDev& m = s->m;
    GP_TRY(sampler_reserve_samples(s, s->outer_done + n_outer));
    for (int it = 0; it < n_outer; it++) {
        const int i = s->outer_done;
        GP_TRY(sampler_mh(s, i, s->sweeps_done, m.nMH));
        s->sweeps_done = 0;
        // binary T: logitT is sliced before the U_k in every ESS pass (src/inference.jl:232-237, 292-297, 346-348)
        const bool do_logit = m.binary && s->h_fdef[m.nX].exists;
        if (do_logit && m.nES > 0) GP_TRY(launch_logit_prior(s, 1, i));
        for (int j = 0; j < m.nES; j++) {
            if (do_logit) GP_TRY(launch_ess_logit(s, j, (uint32_t)(i * m.nES + j)));
            for (int k = 0; k < m.nU; k++) GP_TRY(sampler_ess_u(s, k, (uint32_t)(i * m.nES + j)));
        }
        record_kernel<<<m.n_chains, 256, 0, ctx->stream>>>(s->m, s->c, s->samples, i);
        ctx->launches++;
        GP_CUDA(ctx, cudaGetLastError());
        s->outer_done++;
    }
    return GPSLC_OK;
}

// bench/diagnostic stepping: `count` MH sweeps of every chain without ESS or recording; RNG iteration indices keep
// advancing so no stream is reused
int sampler_mh_sweeps(Sampler* s, int count) {
    GP_TRY(sampler_mh(s, s->outer_done, s->sweeps_done, s->sweeps_done + count));
    s->sweeps_done += count;
    return GPSLC_OK;
}

int sampler_ess_pass(Sampler* s, int pass_index) {
    for (int k = 0; k < s->m.nU; k++) GP_TRY(sampler_ess_u(s, k, (uint32_t)(s->outer_done * s->m.nES + pass_index)));
    return GPSLC_OK;
}

int sampler_get_state(Sampler* s, double* dev_out /* [C][stride] device */) {
    record_kernel<<<s->m.n_chains, 256, 0, s->ctx->stream>>>(s->m, s->c, dev_out, 0);
    s->ctx->launches++;
    GP_CUDA(s->ctx, cudaGetLastError());
    return GPSLC_OK;
}

}  // namespace gpslc
