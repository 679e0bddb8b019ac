// extern "C" surface of libgpslc_b200.so (include/gpslc.h).
#include "../../include/gpslc.h"
#include <cstdlib>
#include "capi_util.cuh"

namespace gpslc {
int launch_cov_build(Ctx*, int, int, int, const double*, const double*, size_t, const double*, const double*, const double*, double*);
int launch_chol_logpdf_dense(Ctx*, int, int, const double*, int, const double*, int, double*, double*, double*, int*);
int launch_rbf_logpdf(Ctx*, int, int, int, const double*, size_t, const double*, const double*, const double*, const double*, int,
                      double*, double*, double*, int*);

__global__ void inv_sq_kernel(const double* ls, double* w, size_t n, int square) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) w[i] = square ? 1.0 / (ls[i] * ls[i]) : 1.0 / ls[i];
}
// w = 1 / ls^2 (square = 1) or 1 / ls (square = 0: the pre-scaling factor of the covariance-build kernels)
int launch_inv_sq(Ctx* ctx, const double* ls, double* w, size_t n, int square) {
    if (n == 0) return GPSLC_OK;
    inv_sq_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ls, w, n, square);
    ctx->launches++;
    GP_CUDA(ctx, cudaGetLastError());
    return GPSLC_OK;
}
}  // namespace gpslc

using namespace gpslc;

struct gpslc_ctx { Ctx c; };

extern "C" {

int gpslc_version(void) { return 100; }

int gpslc_create(int device, gpslc_ctx** out) {
    if (!out) return GPSLC_ERR_ARG;
    *out = nullptr;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0 || device < 0 || device >= ndev) return GPSLC_ERR_NO_DEVICE;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return GPSLC_ERR_NO_DEVICE;
    if (prop.major != 10) return GPSLC_ERR_NO_DEVICE;  // sm_100a binary only; no fallback path exists
    if (cudaSetDevice(device) != cudaSuccess) return GPSLC_ERR_CUDA;
    gpslc_ctx* h = new gpslc_ctx();
    h->c.device = device;
    h->c.num_sms = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking) != cudaSuccess) { delete h; return GPSLC_ERR_CUDA; }
    *out = h;
    return GPSLC_OK;
}

void gpslc_destroy(gpslc_ctx* h) {
    if (!h) return;
    cudaSetDevice(h->c.device);
    cudaStreamSynchronize(h->c.stream);
    if (h->c.scratch) cudaFree(h->c.scratch);
    if (h->c.zbuf) cudaFree(h->c.zbuf);
    if (h->c.counter) cudaFree(h->c.counter);
    if (h->c.arena) cudaFree(h->c.arena);
    h->c.block_cache_release();
    cudaStreamDestroy(h->c.stream);
    delete h;
}

const char* gpslc_last_error(const gpslc_ctx* h) { return h ? h->c.last_error.c_str() : "null context"; }

int gpslc_synchronize(gpslc_ctx* h) {
    if (!h) return GPSLC_ERR_ARG;
    GP_CUDA(&h->c, cudaStreamSynchronize(h->c.stream));
    return GPSLC_OK;
}
void* gpslc_stream(gpslc_ctx* h) { return h ? (void*)h->c.stream : nullptr; }
unsigned long long gpslc_launch_count(const gpslc_ctx* h) { return h ? h->c.launches : 0ull; }

int gpslc_malloc(gpslc_ctx* h, size_t bytes, void** dptr) {
    if (!h || !dptr) return GPSLC_ERR_ARG;
    GP_CUDA(&h->c, cudaSetDevice(h->c.device));
    GP_CUDA(&h->c, cudaMalloc(dptr, bytes));
    return GPSLC_OK;
}
int gpslc_free(gpslc_ctx* h, void* dptr) {
    if (!h) return GPSLC_ERR_ARG;
    GP_CUDA(&h->c, cudaFree(dptr));
    return GPSLC_OK;
}
int gpslc_memcpy_h2d(gpslc_ctx* h, void* dst, const void* src, size_t bytes) {
    if (!h) return GPSLC_ERR_ARG;
    GP_CUDA(&h->c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, h->c.stream));
    GP_CUDA(&h->c, cudaStreamSynchronize(h->c.stream));
    return GPSLC_OK;
}
int gpslc_memcpy_d2h(gpslc_ctx* h, void* dst, const void* src, size_t bytes) {
    if (!h) return GPSLC_ERR_ARG;
    GP_CUDA(&h->c, cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, h->c.stream));
    GP_CUDA(&h->c, cudaStreamSynchronize(h->c.stream));
    return GPSLC_OK;
}

int gpslc_cov_build(gpslc_ctx* h, int loc, int n, int batch, int D, const double* f1, const double* f2, int feat_shared,
                    const double* ls, const double* scale, const double* noise, double* K) {
    if (!h) return GPSLC_ERR_ARG;
    Ctx* ctx = &h->c;
    if (n < 0 || batch < 0 || D < 0 || D > DMAX || !K && n > 0 && batch > 0) return ctx->fail(GPSLC_ERR_ARG, "gpslc_cov_build: bad argument");
    if (n == 0 || batch == 0) return GPSLC_OK;
    if (D > 0 && (!f1 || !f2 || !ls)) return ctx->fail(GPSLC_ERR_ARG, "gpslc_cov_build: null feature/lengthscale pointer");
    if (!scale) return ctx->fail(GPSLC_ERR_ARG, "gpslc_cov_build: null scale");
    GP_CUDA(ctx, cudaSetDevice(ctx->device));
    const size_t nf = (size_t)(feat_shared ? 1 : batch) * D * n;
    ArenaScope arena_scope(ctx);
    Staged<double> df1(ctx), df2(ctx), dls(ctx), dsc(ctx), dnz(ctx), dK(ctx), dw(ctx);
    GP_TRY(df1.in(loc, f1, nf));
    if (f2 == f1) { df2.d = df1.d; } else GP_TRY(df2.in(loc, f2, nf));
    GP_TRY(dls.in(loc, ls, (size_t)batch * D));
    GP_TRY(dsc.in(loc, scale, batch));
    GP_TRY(dnz.in(loc, noise, batch));
    GP_TRY(dK.outbuf(loc, K, (size_t)batch * n * n));
    GP_TRY(dw.outbuf(1, nullptr, 0));
    double* w = nullptr;
    if (D > 0) { GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&w), (size_t)batch * D * sizeof(double))); }
    int rc = launch_inv_sq(ctx, dls.d, w, (size_t)batch * D, 0);
    if (!rc) rc = launch_cov_build(ctx, n, batch, D, df1.d, df2.d, feat_shared ? 0 : (size_t)D * n, w, dsc.d, noise ? dnz.d : nullptr, dK.d);
    if (!rc) rc = dK.finish();
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc) return rc;
    if (e != cudaSuccess) return ctx->cuda_fail(e, "gpslc_cov_build");
    return GPSLC_OK;
}

int gpslc_chol_logpdf(gpslc_ctx* h, int loc, int n, int batch, const double* K, int ld, const double* y, int y_shared,
                      double* logpdf, double* logdet, double* quad, int* info) {
    if (!h) return GPSLC_ERR_ARG;
    Ctx* ctx = &h->c;
    if (n <= 0 || batch < 0 || ld < n || !K || !y) return ctx->fail(GPSLC_ERR_ARG, "gpslc_chol_logpdf: bad argument");
    if (batch == 0) return GPSLC_OK;
    GP_CUDA(ctx, cudaSetDevice(ctx->device));
    ArenaScope arena_scope(ctx);
    Staged<double> dK(ctx), dy(ctx), dlp(ctx), dld(ctx), dq(ctx);
    Staged<int> dinfo(ctx);
    GP_TRY(dK.in(loc, K, (size_t)batch * ld * n));
    GP_TRY(dy.in(loc, y, (size_t)(y_shared ? 1 : batch) * n));
    GP_TRY(dlp.outbuf(loc, logpdf, batch));
    GP_TRY(dld.outbuf(loc, logdet, batch));
    GP_TRY(dq.outbuf(loc, quad, batch));
    GP_TRY(dinfo.outbuf(loc, info, batch));
    GP_TRY(launch_chol_logpdf_dense(ctx, batch, n, dK.d, ld, dy.d, y_shared, dlp.d, dld.d, dq.d, dinfo.d));
    GP_TRY(dlp.finish()); GP_TRY(dld.finish()); GP_TRY(dq.finish()); GP_TRY(dinfo.finish());
    GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GPSLC_OK;
}

int gpslc_rbf_logpdf(gpslc_ctx* h, int loc, int n, int batch, int D, const double* feat, int feat_shared, const double* ls,
                     const double* scale, const double* noise, const double* y, int y_shared, double* logpdf,
                     double* logdet, double* quad, int* info) {
    if (!h) return GPSLC_ERR_ARG;
    Ctx* ctx = &h->c;
    if (n <= 0 || batch < 0 || D < 0 || D > DMAX || !scale || !noise || !y || (D > 0 && (!feat || !ls)))
        return ctx->fail(GPSLC_ERR_ARG, "gpslc_rbf_logpdf: bad argument");
    if (batch == 0) return GPSLC_OK;
    GP_CUDA(ctx, cudaSetDevice(ctx->device));
    ArenaScope arena_scope(ctx);
    Staged<double> df(ctx), dls(ctx), dsc(ctx), dnz(ctx), dy(ctx), dlp(ctx), dld(ctx), dq(ctx);
    Staged<int> dinfo(ctx);
    GP_TRY(df.in(loc, feat, (size_t)(feat_shared ? 1 : batch) * D * n));
    GP_TRY(dls.in(loc, ls, (size_t)batch * D));
    GP_TRY(dsc.in(loc, scale, batch));
    GP_TRY(dnz.in(loc, noise, batch));
    GP_TRY(dy.in(loc, y, (size_t)(y_shared ? 1 : batch) * n));
    GP_TRY(dlp.outbuf(loc, logpdf, batch));
    GP_TRY(dld.outbuf(loc, logdet, batch));
    GP_TRY(dq.outbuf(loc, quad, batch));
    GP_TRY(dinfo.outbuf(loc, info, batch));
    double* w = nullptr;
    if (D > 0) GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&w), (size_t)batch * D * sizeof(double)));
    int rc = launch_inv_sq(ctx, dls.d, w, (size_t)batch * D, 1);
    if (!rc) rc = launch_rbf_logpdf(ctx, batch, n, D, df.d, feat_shared ? 0 : (size_t)D * n, w, dsc.d, dnz.d, dy.d, y_shared,
                                    dlp.d, dld.d, dq.d, dinfo.d);
    if (!rc) rc = dlp.finish();
    if (!rc) rc = dld.finish();
    if (!rc) rc = dq.finish();
    if (!rc) rc = dinfo.finish();
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (rc) return rc;
    if (e != cudaSuccess) return ctx->cuda_fail(e, "gpslc_rbf_logpdf");
    return GPSLC_OK;
}

}  // extern "C"

// ================================================================================================ sampler
#include "sampler.cuh"
namespace gpslc {
int sampler_create(Ctx*, int, int, int, int, int, const double*, const double*, const double*, int, const int*, const double*, double, double,
                   const double*, const double*, double, int, int, int, unsigned long long, int, int, int, int, int, Sampler**);
void sampler_free(Sampler*);
int sampler_init(Sampler*);
int sampler_run(Sampler*, int);
int sampler_mh_sweeps(Sampler*, int);
int sampler_ess_pass(Sampler*, int);
int sampler_get_state(Sampler*, double*);
int sampler_set_state(Sampler*, int, const double*);
}
struct gpslc_sampler { Sampler* s; };

extern "C" {

int gpslc_sampler_create(gpslc_ctx* h, int loc, const gpslc_data* d, const gpslc_prior* p, const gpslc_opts* o, gpslc_sampler** out) {
    if (!h || !out) return GPSLC_ERR_ARG;
    Ctx* ctx = &h->c;
    *out = nullptr;
    if (!d || !p || !o) return ctx->fail(GPSLC_ERR_ARG, "gpslc_sampler_create: null argument");
    if (o->nOuter < 0 || o->nMHInner < 0 || o->nESInner < 0) return ctx->fail(GPSLC_ERR_ARG, "gpslc_sampler_create: negative iteration count");
    Sampler* s = nullptr;
    GP_TRY(sampler_create(ctx, loc, d->n, d->nX, d->nU, d->binary, d->X, d->T, d->Y, d->n_obj, d->obj_counts, d->sigma_u_dense, d->sigma_u_eps,
                          d->sigma_u_cov, p->shape, p->scale, p->drift, o->nMHInner, o->nESInner, o->n_chains, o->seed,
                          o->chain_offset, o->u_layout_mode, o->ess_rule, o->observe_x, d->per_chain_data, &s));
    int rc = sampler_init(s);
    if (rc) { sampler_free(s); return rc; }
    *out = new gpslc_sampler{s};
    return GPSLC_OK;
}

void gpslc_sampler_destroy(gpslc_sampler* hs) {
    if (!hs) return;
    sampler_free(hs->s);
    delete hs;
}

int gpslc_sampler_layout(const gpslc_sampler* hs, int* n_params, int* stride, int* n_sites, int* n_factors) {
    if (!hs) return GPSLC_ERR_ARG;
    if (n_params) *n_params = hs->s->m.n_params;
    if (stride) *stride = hs->s->m.stride;
    if (n_sites) *n_sites = hs->s->m.n_sites;
    if (n_factors) *n_factors = hs->s->m.nF;
    return GPSLC_OK;
}

int gpslc_sampler_run(gpslc_sampler* hs, int n_outer) {
    if (!hs || n_outer < 0) return GPSLC_ERR_ARG;
    GP_CUDA(hs->s->ctx, cudaSetDevice(hs->s->ctx->device));
    return sampler_run(hs->s, n_outer);
}
int gpslc_sampler_mh_sweeps(gpslc_sampler* hs, int count) {
    if (!hs || count < 0) return GPSLC_ERR_ARG;
    GP_CUDA(hs->s->ctx, cudaSetDevice(hs->s->ctx->device));
    return sampler_mh_sweeps(hs->s, count);
}
int gpslc_sampler_ess_pass(gpslc_sampler* hs, int pass_index) {
    if (!hs) return GPSLC_ERR_ARG;
    GP_CUDA(hs->s->ctx, cudaSetDevice(hs->s->ctx->device));
    return sampler_ess_pass(hs->s, pass_index);
}

int gpslc_sampler_get_samples(gpslc_sampler* hs, int loc, double* out, int* outer_done) {
    if (!hs) return GPSLC_ERR_ARG;
    Sampler* s = hs->s; Ctx* ctx = s->ctx;
    if (outer_done) *outer_done = s->outer_done;
    if (out && s->outer_done > 0) {
        const size_t bytes = (size_t)s->outer_done * s->m.n_chains * s->m.stride * sizeof(double);
        GP_CUDA(ctx, cudaMemcpyAsync(out, s->samples, bytes, loc == 1 ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, ctx->stream));
    }
    GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GPSLC_OK;
}
const double* gpslc_sampler_samples_device(gpslc_sampler* hs) { return hs ? hs->s->samples : nullptr; }

int gpslc_sampler_get_state(gpslc_sampler* hs, int loc, double* packed) {
    if (!hs || !packed) return GPSLC_ERR_ARG;
    Sampler* s = hs->s; Ctx* ctx = s->ctx;
    const size_t count = (size_t)s->m.n_chains * s->m.stride;
    if (loc == 1) { GP_TRY(sampler_get_state(s, packed)); GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream)); return GPSLC_OK; }
    double* tmp = nullptr;
    GP_CUDA(ctx, cudaMalloc(&tmp, count * sizeof(double)));
    int rc = sampler_get_state(s, tmp);
    cudaError_t e = cudaMemcpyAsync(packed, tmp, count * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream);
    cudaError_t e2 = cudaStreamSynchronize(ctx->stream);
    cudaFree(tmp);
    if (rc) return rc;
    if (e != cudaSuccess) return ctx->cuda_fail(e, "gpslc_sampler_get_state");
    if (e2 != cudaSuccess) return ctx->cuda_fail(e2, "gpslc_sampler_get_state");
    return GPSLC_OK;
}
int gpslc_sampler_set_state(gpslc_sampler* hs, int loc, const double* packed) {
    if (!hs || !packed) return GPSLC_ERR_ARG;
    return sampler_set_state(hs->s, loc, packed);
}
int gpslc_sampler_get_terms(gpslc_sampler* hs, double* factor_logpdf, double* u_quad) {
    if (!hs) return GPSLC_ERR_ARG;
    Sampler* s = hs->s; Ctx* ctx = s->ctx;
    if (factor_logpdf) GP_CUDA(ctx, cudaMemcpyAsync(factor_logpdf, s->c.lp, (size_t)s->m.n_chains * s->m.nF * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (u_quad && s->m.nU > 0) GP_CUDA(ctx, cudaMemcpyAsync(u_quad, s->c.q, (size_t)s->m.n_chains * s->m.nU * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GPSLC_OK;
}
int gpslc_sampler_get_stats(gpslc_sampler* hs, unsigned long long* accepts, unsigned long long* ess_evals, unsigned long long* ess_evals_logit) {
    if (!hs) return GPSLC_ERR_ARG;
    Sampler* s = hs->s; Ctx* ctx = s->ctx;
    if (accepts) GP_CUDA(ctx, cudaMemcpyAsync(accepts, s->c.accepts, (size_t)s->m.n_chains * s->m.n_sites * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (ess_evals) GP_CUDA(ctx, cudaMemcpyAsync(ess_evals, s->c.ess_evals, (size_t)s->m.n_chains * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    if (ess_evals_logit) GP_CUDA(ctx, cudaMemcpyAsync(ess_evals_logit, s->c.ess_evals_logit, (size_t)s->m.n_chains * sizeof(unsigned long long), cudaMemcpyDeviceToHost, ctx->stream));
    GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GPSLC_OK;
}

int gpslc_posterior(gpslc_ctx* h, const gpslc_data* d, const gpslc_prior* p, const gpslc_opts* o, double* samples_out,
                    unsigned long long* accepts, unsigned long long* ess_evals) {
    if (!h || !o) return GPSLC_ERR_ARG;
    gpslc_sampler* hs = nullptr;
    GP_TRY(gpslc_sampler_create(h, GPSLC_HOST, d, p, o, &hs));
    int rc = gpslc_sampler_run(hs, o->nOuter);
    if (!rc) rc = gpslc_sampler_get_samples(hs, GPSLC_HOST, samples_out, nullptr);
    if (!rc) rc = gpslc_sampler_get_stats(hs, accepts, ess_evals, nullptr);
    gpslc_sampler_destroy(hs);
    return rc;
}

}  // extern "C"

// ================================================================================================ estimation
#include "est.cuh"
namespace gpslc {
int launch_ite(Ctx*, const EstArgs&);
int launch_sate(Ctx*, const EstArgs&);
}

static int est_common(gpslc_ctx* h, int loc, const gpslc_data* d, const double* samples, int n_outer, int n_chains, int stride,
                      const int* ret_idx, int R, const double* doT, int n_doT, double jitter, int spp, uint64_t seed,
                      int chain_offset, int var_as_std, bool sate, double* o1, double* o2, double* o3, int* info, int dot_offset = 0,
                      double summary_ci = 0.0, double* summary = nullptr, const unsigned char* subset_mask = nullptr,
                      double* subset_sate = nullptr) {
    if (!h) return GPSLC_ERR_ARG;
    Ctx* ctx = &h->c;
    if (!d || !samples || !ret_idx || !doT || d->n <= 0 || R < 0 || n_doT < 0 || n_chains <= 0 || spp < 0)
        return ctx->fail(GPSLC_ERR_ARG, "gpslc_ite/sate: bad argument");
    const int n = d->n, nX = d->nX, nU = d->nU;
    const int n_params = 6 + 4 * nX + 2 * nU + nU * nX;
    if (stride < n_params + nU * n) return ctx->fail(GPSLC_ERR_ARG, "gpslc_ite/sate: stride too small for the packed layout");
    if (nU + nX + 1 > DMAX) return ctx->fail(GPSLC_ERR_UNSUPPORTED, "gpslc_ite/sate: nU + nX + 1 exceeds DMAX");
    if (R == 0 || n_doT == 0) return GPSLC_OK;
    GP_CUDA(ctx, cudaSetDevice(ctx->device));
    // ret_idx / doT are tiny host arrays in either mode
    for (int r = 0; r < R; r++)
        if (ret_idx[r] < 0 || ret_idx[r] >= n_outer) return ctx->fail(GPSLC_ERR_ARG, "gpslc_ite/sate: retained index out of range");
    ArenaScope arena_scope(ctx);
    Staged<double> dX(ctx), dT(ctx), dY(ctx), dS(ctx), dDo(ctx), d1(ctx), d2(ctx), d3(ctx);
    Staged<int> dRet(ctx), dInfo(ctx);
    GP_TRY(dX.in(loc, d->X, (size_t)n * nX));
    GP_TRY(dT.in(loc, d->T, n));
    GP_TRY(dY.in(loc, d->Y, n));
    GP_TRY(dS.in(loc, samples, (size_t)n_outer * n_chains * stride));
    GP_TRY(dRet.in(0, ret_idx, R));
    GP_TRY(dDo.in(0, doT, n_doT));
    const size_t tasks = (size_t)n_doT * n_chains * R;
    EstArgs a{};
    a.n = n; a.nX = nX; a.nU = nU; a.n_params = n_params; a.stride = stride;
    a.X = dX.d; a.T = dT.d; a.Y = dY.d; a.samples = dS.d; a.n_chains = n_chains; a.ret_idx = dRet.d; a.R = R;
    a.doT = dDo.d; a.n_doT = n_doT; a.jitter = jitter; a.spp = spp; a.seed = seed; a.chain0 = chain_offset;
    a.var_as_std = var_as_std; a.dot0 = dot_offset;
    { const char* e = getenv("GPSLC_LS_UNSQUARED"); a.ls_unsquared = (e && atoi(e) == 1) ? 1 : 0; }
    GP_TRY(dInfo.outbuf(loc, info, tasks));
    a.info = dInfo.d;
    int rc;
    Staged<double> dSum(ctx), dSub(ctx);
    if (!sate && summary) {
        // fused predictCounterfactualEffects + summarizeEstimates: the draws live only in the library's device arena
        if (spp <= 0) return ctx->fail(GPSLC_ERR_ARG, "gpslc_ite_summary: samplesPerPosterior must be positive");
        double* draws = nullptr;
        GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&draws), tasks * spp * n * sizeof(double)));
        a.mean_out = nullptr; a.cov_out = nullptr; a.ite_out = draws;
        if (!subset_mask) {
            GP_TRY(dSum.outbuf(loc, summary, (size_t)n_doT * n_chains * n * 3));
            rc = launch_ite(ctx, a);
            if (!rc) rc = launch_summarize(ctx, draws, n_doT * n_chains, R * spp, n, summary_ci, dSum.d);
        } else {
            // subgroup effect curve (docs/src/index.md:101-114): draws -> mean over the subset -> per (chain, doT) summary; the mask
            // is a small host array in either mode
            int cnt = 0;
            for (int i = 0; i < n; i++) cnt += subset_mask[i] != 0;
            if (cnt == 0) return ctx->fail(GPSLC_ERR_ARG, "gpslc_ite_subset_summary: the subset is empty");
            Staged<unsigned char> dMask(ctx);
            GP_TRY(dMask.in(0, subset_mask, n));
            const size_t nsub = (size_t)n_chains * R * spp * n_doT;
            GP_TRY(dSub.outbuf(loc, subset_sate, nsub));
            double* sub = dSub.d;
            if (!sub) GP_CUDA(ctx, ctx->arena_alloc(reinterpret_cast<void**>(&sub), nsub * sizeof(double)));
            GP_TRY(dSum.outbuf(loc, summary, (size_t)n_chains * n_doT * 3));
            rc = launch_ite(ctx, a);
            if (!rc) rc = launch_subset_mean(ctx, draws, dMask.d, tasks * spp, R * spp, n, n_chains, n_doT, cnt, sub);
            if (!rc) rc = launch_summarize(ctx, sub, n_chains, R * spp, n_doT, summary_ci, dSum.d);
            if (!rc) rc = dSub.finish();
        }
        if (!rc) rc = dSum.finish();
    } else if (!sate) {
        GP_TRY(d1.outbuf(loc, o1, tasks * n));
        GP_TRY(d2.outbuf(loc, o2, tasks * n * n));
        GP_TRY(d3.outbuf(loc, o3, tasks * spp * n));
        a.mean_out = d1.d; a.cov_out = d2.d; a.ite_out = d3.d;
        rc = launch_ite(ctx, a);
    } else {
        GP_TRY(d1.outbuf(loc, o1, tasks));
        GP_TRY(d2.outbuf(loc, o2, tasks));
        GP_TRY(d3.outbuf(loc, o3, tasks * spp));
        a.msate = d1.d; a.vsate = d2.d; a.sate_out = d3.d;
        rc = launch_sate(ctx, a);
    }
    if (rc) return rc;
    GP_TRY(d1.finish()); GP_TRY(d2.finish()); GP_TRY(d3.finish()); GP_TRY(dInfo.finish());
    GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GPSLC_OK;
}

extern "C" {
int gpslc_ite(gpslc_ctx* h, int loc, const gpslc_data* d, const double* samples, int n_outer, int n_chains, int stride,
              const int* ret_idx, int R, const double* doT, int n_doT, double jitter, int spp, uint64_t seed, int chain_offset,
              double* meanITE, double* covITE, double* ite, int* info) {
    return est_common(h, loc, d, samples, n_outer, n_chains, stride, ret_idx, R, doT, n_doT, jitter, spp, seed, chain_offset, 0,
                      false, meanITE, covITE, ite, info);
}
int gpslc_ite_slice(gpslc_ctx* h, int loc, const gpslc_data* d, const double* samples, int n_outer, int n_chains, int stride,
                    const int* ret_idx, int R, const double* doT, int n_doT, int dot_offset, double jitter, int spp, uint64_t seed,
                    int chain_offset, double* meanITE, double* covITE, double* ite, int* info) {
    if (dot_offset < 0) return GPSLC_ERR_ARG;
    return est_common(h, loc, d, samples, n_outer, n_chains, stride, ret_idx, R, doT, n_doT, jitter, spp, seed, chain_offset, 0,
                      false, meanITE, covITE, ite, info, dot_offset);
}
int gpslc_ite_summary(gpslc_ctx* h, int loc, const gpslc_data* d, const double* samples, int n_outer, int n_chains, int stride,
                      const int* ret_idx, int R, const double* doT, int n_doT, int dot_offset, double jitter, int spp, uint64_t seed,
                      int chain_offset, double credible_interval, double* summary, int* info) {
    if (dot_offset < 0 || !summary || !(credible_interval > 0.0 && credible_interval < 1.0)) return GPSLC_ERR_ARG;
    return est_common(h, loc, d, samples, n_outer, n_chains, stride, ret_idx, R, doT, n_doT, jitter, spp, seed, chain_offset, 0,
                      false, nullptr, nullptr, nullptr, info, dot_offset, credible_interval, summary);
}
int gpslc_ite_subset_summary(gpslc_ctx* h, int loc, const gpslc_data* d, const double* samples, int n_outer, int n_chains, int stride,
                             const int* ret_idx, int R, const double* doT, int n_doT, int dot_offset, double jitter, int spp,
                             uint64_t seed, int chain_offset, const unsigned char* mask, double credible_interval, double* sate,
                             double* summary, int* info) {
    if (dot_offset < 0 || !summary || !mask || !(credible_interval > 0.0 && credible_interval < 1.0)) return GPSLC_ERR_ARG;
    return est_common(h, loc, d, samples, n_outer, n_chains, stride, ret_idx, R, doT, n_doT, jitter, spp, seed, chain_offset, 0,
                      false, nullptr, nullptr, nullptr, info, dot_offset, credible_interval, summary, mask, sate);
}
int gpslc_subset_mean(gpslc_ctx* h, int loc, const double* samples, int batch, int m, int n, const unsigned char* mask, double* out) {
    if (!h) return GPSLC_ERR_ARG;
    Ctx* ctx = &h->c;
    if (!samples || !out || !mask || batch < 0 || m <= 0 || n <= 0) return ctx->fail(GPSLC_ERR_ARG, "gpslc_subset_mean: bad argument");
    int cnt = 0;
    for (int i = 0; i < n; i++) cnt += mask[i] != 0;
    if (cnt == 0) return ctx->fail(GPSLC_ERR_ARG, "gpslc_subset_mean: the subset is empty");
    if (batch == 0) return GPSLC_OK;
    GP_CUDA(ctx, cudaSetDevice(ctx->device));
    ArenaScope arena_scope(ctx);
    Staged<double> dS(ctx), dO(ctx);
    Staged<unsigned char> dM(ctx);
    GP_TRY(dS.in(loc, samples, (size_t)batch * m * n));
    GP_TRY(dM.in(0, mask, n));
    GP_TRY(dO.outbuf(loc, out, (size_t)batch * m));
    GP_TRY(launch_subset_mean(ctx, dS.d, dM.d, (size_t)batch * m, m, n, 0, 0, cnt, dO.d));
    GP_TRY(dO.finish());
    GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GPSLC_OK;
}
int gpslc_sate_slice(gpslc_ctx* h, int loc, const gpslc_data* d, const double* samples, int n_outer, int n_chains, int stride,
                     const int* ret_idx, int R, const double* doT, int n_doT, int dot_offset, double jitter, int spp, uint64_t seed,
                     int chain_offset, int var_as_std, double* meanSATE, double* varSATE, double* sate, int* info) {
    if (dot_offset < 0) return GPSLC_ERR_ARG;
    return est_common(h, loc, d, samples, n_outer, n_chains, stride, ret_idx, R, doT, n_doT, jitter, spp, seed, chain_offset,
                      var_as_std, true, meanSATE, varSATE, sate, info, dot_offset);
}
int gpslc_summarize(gpslc_ctx* h, int loc, const double* samples, int batch, int m, int n, double credible_interval, double* out) {
    if (!h) return GPSLC_ERR_ARG;
    Ctx* ctx = &h->c;
    if (!samples || !out || batch < 0 || m <= 0 || n <= 0 || !(credible_interval > 0.0 && credible_interval < 1.0))
        return ctx->fail(GPSLC_ERR_ARG, "gpslc_summarize: bad argument");
    if (batch == 0) return GPSLC_OK;
    GP_CUDA(ctx, cudaSetDevice(ctx->device));
    ArenaScope arena_scope(ctx);
    Staged<double> dS(ctx), dO(ctx);
    GP_TRY(dS.in(loc, samples, (size_t)batch * m * n));
    GP_TRY(dO.outbuf(loc, out, (size_t)batch * n * 3));
    GP_TRY(launch_summarize(ctx, dS.d, batch, m, n, credible_interval, dO.d));
    GP_TRY(dO.finish());
    GP_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return GPSLC_OK;
}
int gpslc_sate(gpslc_ctx* h, int loc, const gpslc_data* d, const double* samples, int n_outer, int n_chains, int stride,
               const int* ret_idx, int R, const double* doT, int n_doT, double jitter, int spp, uint64_t seed, int chain_offset,
               int var_as_std, double* meanSATE, double* varSATE, double* sate, int* info) {
    return est_common(h, loc, d, samples, n_outer, n_chains, stride, ret_idx, R, doT, n_doT, jitter, spp, seed, chain_offset,
                      var_as_std, true, meanSATE, varSATE, sate, info);
}
}
