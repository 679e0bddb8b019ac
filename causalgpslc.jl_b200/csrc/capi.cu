// extern "C" surface of libgpslc_b200.so (include/gpslc.h).
#include "../../include/gpslc.h"
#include "capi_util.cuh"

namespace gpslc {
int launch_cov_build(Ctx*, int, int, int, const double*, const double*, size_t, const double*, const double*, const double*, double*);
int launch_chol_logpdf_dense(Ctx*, int, int, const double*, int, const double*, int, double*, double*, double*, int*);
int launch_rbf_logpdf(Ctx*, int, int, int, const double*, size_t, const double*, const double*, const double*, const double*, int,
                      double*, double*, double*, int*);

__global__ void inv_sq_kernel(const double* ls, double* w, size_t n) {
    size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i < n) w[i] = 1.0 / (ls[i] * ls[i]);
}
int launch_inv_sq(Ctx* ctx, const double* ls, double* w, size_t n) {
    if (n == 0) return GPSLC_OK;
    inv_sq_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ls, w, n);
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
    Staged<double> df1(ctx), df2(ctx), dls(ctx), dsc(ctx), dnz(ctx), dK(ctx), dw(ctx);
    GP_TRY(df1.in(loc, f1, nf));
    if (f2 == f1) { df2.d = df1.d; } else GP_TRY(df2.in(loc, f2, nf));
    GP_TRY(dls.in(loc, ls, (size_t)batch * D));
    GP_TRY(dsc.in(loc, scale, batch));
    GP_TRY(dnz.in(loc, noise, batch));
    GP_TRY(dK.outbuf(loc, K, (size_t)batch * n * n));
    GP_TRY(dw.outbuf(1, nullptr, 0));
    double* w = nullptr;
    if (D > 0) { GP_CUDA(ctx, cudaMalloc(&w, (size_t)batch * D * sizeof(double))); }
    int rc = launch_inv_sq(ctx, dls.d, w, (size_t)batch * D);
    if (!rc) rc = launch_cov_build(ctx, n, batch, D, df1.d, df2.d, feat_shared ? 0 : (size_t)D * n, w, dsc.d, noise ? dnz.d : nullptr, dK.d);
    if (!rc) rc = dK.finish();
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (w) cudaFree(w);
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
    if (D > 0) GP_CUDA(ctx, cudaMalloc(&w, (size_t)batch * D * sizeof(double)));
    int rc = launch_inv_sq(ctx, dls.d, w, (size_t)batch * D);
    if (!rc) rc = launch_rbf_logpdf(ctx, batch, n, D, df.d, feat_shared ? 0 : (size_t)D * n, w, dsc.d, dnz.d, dy.d, y_shared,
                                    dlp.d, dld.d, dq.d, dinfo.d);
    if (!rc) rc = dlp.finish();
    if (!rc) rc = dld.finish();
    if (!rc) rc = dq.finish();
    if (!rc) rc = dinfo.finish();
    cudaError_t e = cudaStreamSynchronize(ctx->stream);
    if (w) cudaFree(w);
    if (rc) return rc;
    if (e != cudaSuccess) return ctx->cuda_fail(e, "gpslc_rbf_logpdf");
    return GPSLC_OK;
}

}  // extern "C"
