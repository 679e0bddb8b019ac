/* gpslc.h — C ABI of libgpslc_b200.so: the B200 (sm_100a) implementation of the GP-SLC hot path of
 * KDL-umass/CausalGPSLC.jl (covariance build -> Cholesky -> MvNormal log-density inside the MCMC over U and the
 * hyperparameters, and the GP-conditional ITE/SATE sampling).
 *
 * The reference has no FFI layer: its seam is a set of Julia functions. Each entry point below names
 * the reference function(s) whose body a Julia maintainer re-points at it with `ccall` (see INTEGRATION.md and
 * causalgpslc.jl_b200/julia/CausalGPSLCB200.jl), citing file:line in /root/reference.
 *
 * Conventions
 *  - plain C types only; all matrices column-major FP64 (Julia `Matrix{Float64}` memory); Bool treatments are passed
 *    as 0.0/1.0 doubles (the reference's kernel subtracts Bools as integers, src/kernel.jl:17 — same values).
 *  - the caller owns every buffer it passes; the library never retains or frees caller memory.
 *  - `loc` says where the caller's buffers live: GPSLC_HOST (the library stages them through its own device buffers,
 *    the normal `ccall` case) or GPSLC_DEVICE (already resident in HBM on ctx's device).
 *  - return value: 0 on success, else a GPSLC_ERR_* code; gpslc_last_error() gives the text. Numerical failure of
 *    one batch element is reported LAPACK-style in `info[]` (0 ok, k>0: leading minor k not positive definite) —
 *    the analogue of the PosDefException the reference lets escape from `cholesky`.
 *  - a ctx is bound to one GPU, owns one stream and its workspaces, and is NOT thread-safe. There is no CPU fallback:
 *    gpslc_create fails with GPSLC_ERR_NO_DEVICE when no sm_100 device is present.
 */
#ifndef GPSLC_H
#define GPSLC_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPSLC_OK 0
#define GPSLC_ERR_CUDA 1
#define GPSLC_ERR_ARG 2
#define GPSLC_ERR_NOT_PD 3
#define GPSLC_ERR_UNSUPPORTED 4
#define GPSLC_ERR_NO_DEVICE 5

#define GPSLC_HOST 0
#define GPSLC_DEVICE 1

typedef struct gpslc_ctx gpslc_ctx;

/* ---- context ------------------------------------------------------------------------------------------------ */
int gpslc_version(void);
int gpslc_create(int device, gpslc_ctx** out);
void gpslc_destroy(gpslc_ctx* ctx);
const char* gpslc_last_error(const gpslc_ctx* ctx);
int gpslc_synchronize(gpslc_ctx* ctx);
/* the ctx's cudaStream_t (as void*), so a harness can bracket calls with CUDA events on the launching stream */
void* gpslc_stream(gpslc_ctx* ctx);
/* number of kernels this ctx has launched so far (bench.py's "gpu_launches") */
unsigned long long gpslc_launch_count(const gpslc_ctx* ctx);
/* device memory helpers for hosts without a CUDA binding of their own */
int gpslc_malloc(gpslc_ctx* ctx, size_t bytes, void** dptr);
int gpslc_free(gpslc_ctx* ctx, void* dptr);
int gpslc_memcpy_h2d(gpslc_ctx* ctx, void* dst, const void* src, size_t bytes);
int gpslc_memcpy_d2h(gpslc_ctx* ctx, void* dst, const void* src, size_t bytes);

/* ---- L1/L0 deterministic primitives (parity layer) ------------------------------------------------------------ */

/* Batched covariance build. Replaces rbfKernelLog (src/kernel.jl:24-42) summed over feature groups followed by
 * processCov (src/kernel.jl:53-59):
 *   K_b[i,j] = scale_b * exp(-sum_d (f1_b[i,d]-f2_b[j,d])^2 / ls_b[d]^2) (+ noise_b * [i==j] when noise != NULL)
 * f1,f2: feature matrices n x D column-major, one per batch element, or a single shared one (feat_shared=1);
 * f2 == f1 for the usual K(X,X); f2 != f1 for K(T, doT) (src/likelihood.jl:27). ls: batch x D; scale, noise: batch.
 * K: batch matrices n x n, column-major, contiguous. HBM-bound: 8 n^2 bytes written per matrix. */
int gpslc_cov_build(gpslc_ctx* ctx, int loc, int n, int batch, int D, const double* f1, const double* f2,
                    int feat_shared, const double* ls, const double* scale, const double* noise, double* K);

/* Batched log N(y; 0, K) for dense K. Replaces Distributions.logpdf(MvNormal(zeros(n), K), y) reached through Gen
 * `mvnormal` (src/model_likelihood.jl:30,41,49,58,68,78,89,99,109,118; src/model_prior.jl:29,35): LAPACK dpotrf +
 * dtrsv + log-diagonal sum. K: batch x (n x n) column-major with leading dimension ld (lower triangle referenced);
 * y: batch x n, or one shared vector (y_shared=1). Outputs (each may be NULL): logpdf, logdet (= log det K),
 * quad (= y' K^-1 y), info. */
int gpslc_chol_logpdf(gpslc_ctx* ctx, int loc, int n, int batch, const double* K, int ld, const double* y, int y_shared,
                      double* logpdf, double* logdet, double* quad, int* info);

/* Fused build + Cholesky + log-density: the covariance of gpslc_cov_build (f2 == f1, with noise) is generated
 * inside the factorisation tiles and never written to HBM. This is the kernel the sampler uses for every GP factor
 * of src/model_likelihood.jl:13-120. Same outputs as gpslc_chol_logpdf. */
int gpslc_rbf_logpdf(gpslc_ctx* ctx, int loc, int n, int batch, int D, const double* feat, int feat_shared,
                     const double* ls, const double* scale, const double* noise, const double* y, int y_shared,
                     double* logpdf, double* logdet, double* quad, int* info);

#ifdef __cplusplus
}
#endif
#endif /* GPSLC_H */
