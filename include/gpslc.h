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

/* ---- L3: the many-chain posterior sampler ------------------------------------------------------------------- */

/* Observed data + confounder structure: the fields SigmaU, obj, X, T, Y of GPSLCObject (src/types.jl:249-258).
 * SigmaU is passed as the object counts generateSigmaU (src/utils.jl:17-33) was built from, plus its eps/cov
 * (priorparams["sigmaUNoise"], ["sigmaUCov"]); the library uses the closed form of that block matrix. Any other SigmaU
 * goes in densely through sigma_u_dense. */
typedef struct {
    int n;                 /* individuals */
    int nX;                /* covariates; 0 == `X === nothing` */
    int nU;                /* latent confounder dimensions (hyperparams.nU); 0 == `nothing` (no SigmaU) */
    int binary;            /* 1 when T is Vector{Bool} */
    const double* X;       /* n x nX column-major, or NULL */
    const double* T;       /* n (0.0/1.0 when binary) */
    const double* Y;       /* n */
    int n_obj;             /* number of objects (blocks of SigmaU) */
    const int* obj_counts; /* their sizes, in row order (rows sorted by obj as prepareData does, src/data.jl:25) */
    double sigma_u_eps;    /* 1e-13 by default (src/hyperparameters.jl:66) */
    double sigma_u_cov;    /* 1.0 by default  (src/hyperparameters.jl:67) */
    int per_chain_data;    /* 0: X, T, Y are one dataset shared by all chains (the reference's case); 1: X, T, Y hold n_chains
                              datasets back to back (chain-major) — simulation-based calibration runs, test/sbc.jl shape */
    const double* sigma_u_dense; /* NULL, or (with n_obj == 0) an arbitrary symmetric positive definite SigmaU, n x n column-major:
                              samplePosterior(hyperparams, priorparams, SigmaU, X, T, Y) accepts any matrix (src/driver.jl:59-69)
                              and generateU factors uNoise*SigmaU on every update (src/model_prior.jl:27-30). The library factors
                              SigmaU once (GPSLC_ERR_NOT_PD if that fails) and uses L_S for the U prior density and draws. */
} gpslc_data;

/* InvGamma(shape, scale) priors of src/hyperparameters.jl:39-65 in the order
 * uNoise, xNoise, tNoise, yNoise, xScale, tScale, yScale, uxLS, utLS, xtLS, uyLS, xyLS, tyLS; and the MH drift. */
typedef struct {
    double shape[13];
    double scale[13];
    double drift;
} gpslc_prior;

typedef struct {
    int nOuter, nMHInner, nESInner; /* HyperParameters (src/types.jl:22-30) */
    int n_chains;                   /* independent chains run by this call (the reference runs 1) */
    uint64_t seed;                  /* Philox key */
    int chain_offset;               /* global id of chain 0 (multi-GPU sharding keeps streams independent of the split) */
    int u_layout_mode;              /* 0: reference toMatrix interleave (identity when nU == 1); 1: column-wise */
    int ess_rule;                   /* acceptance test of every elliptical_slice update (U_k and, for binary T, logitT):
                                       0: Gen's `while weight <= log(u)` with the FULL `update` weight, i.e. including the Gaussian
                                          prior density of the sliced address (Gen 0.4 inference/elliptical_slice.jl as recollected,
                                          SURVEY.md App. C). Default, because the target is the reference's behaviour. CAVEAT: this
                                          rule counts the prior twice, so the chain does NOT target the model's posterior:
                                          simulation-based calibration fails for uNoise and the U-dependent statistics
                                          (p ~ 0, profiles/sbc_r01.md; asserted in tests/test_gpu_sbc.py).
                                       1: textbook elliptical slice sampling (likelihood terms only): calibrated. */
    int observe_x;                  /* no-U models: 0 = reference behaviour (X never observed, App. B3), 1 = condition on X */
} gpslc_opts;

typedef struct gpslc_sampler gpslc_sampler;

/* Builds the model tables, uploads the data and performs Gen `generate` (src/inference.jl:20,73,123,156,189,261,321,370):
 * every chain's hyperparameters are drawn from their priors and U_k ~ N(0, uNoise*SigmaU). Fails with
 * GPSLC_ERR_NOT_PD if an initial covariance is not positive definite (the reference throws PosDefException). */
int gpslc_sampler_create(gpslc_ctx* ctx, int loc, const gpslc_data* data, const gpslc_prior* prior, const gpslc_opts* opts,
                         gpslc_sampler** out);
void gpslc_sampler_destroy(gpslc_sampler* s);
/* packed sample layout: [uNoise tNoise yNoise tyLS tScale yScale | xNoise[nX] | xScale[nX] | xtLS[nX] |
 * xyLS[nX] | utLS[nU] | uyLS[nU] | uxLS[nU*nX, i-major] | U[nU*n, i-major] | logitT[n] if binary | Xmodel[n*nX] if any];
 * unused hyperparameters are NaN. n_params = length of the hyperparameter part; stride = record length. */
int gpslc_sampler_layout(const gpslc_sampler* s, int* n_params, int* stride, int* n_sites, int* n_factors);
/* Run n_outer more outer iterations of `Posterior` (src/inference.jl:21-57): nMHInner MH sweeps, nESInner elliptical
 * slice passes (logitT first when T is binary, then U_1..U_nU), then record the sample. */
int gpslc_sampler_run(gpslc_sampler* s, int n_outer);
/* Stepping entry points used by bench.py: `count` MH sweeps of every chain (one sweep = the body of the
 * `for j = 1:nMHInner` loop, src/inference.jl:22-45) / one ESS pass over all U_k. */
int gpslc_sampler_mh_sweeps(gpslc_sampler* s, int count);
int gpslc_sampler_ess_pass(gpslc_sampler* s, int pass_index);
/* samples recorded so far: [outer_done][n_chains][stride] */
int gpslc_sampler_get_samples(gpslc_sampler* s, int loc, double* out, int* outer_done);
/* device pointer of the same buffer (valid until the next gpslc_sampler_run), for consumers that stay on the GPU */
const double* gpslc_sampler_samples_device(gpslc_sampler* s);
int gpslc_sampler_get_state(gpslc_sampler* s, int loc, double* packed /* [n_chains][stride] */);
int gpslc_sampler_set_state(gpslc_sampler* s, int loc, const double* packed);
/* cached log-joint pieces of the current state: log N(.) of every GP factor [n_chains][nX+2] (X_1..X_nX, T, Y; entries of
 * factors that do not exist in the model variant are 0) and U_k' SigmaU^-1 U_k [n_chains][nU]. Parity layer for
 * the log-joint of src/model.jl:11-130. */
int gpslc_sampler_get_terms(gpslc_sampler* s, double* factor_logpdf, double* u_quad);
/* accepted MH moves per (chain, site) in sweep order; elliptical-slice model evaluations per chain for the U_k updates and
 * (binary T) for the logitT updates. Each pointer may be NULL. */
int gpslc_sampler_get_stats(gpslc_sampler* s, unsigned long long* accepts, unsigned long long* ess_evals,
                            unsigned long long* ess_evals_logit);

/* One-shot form, the body of `Posterior(priorparams, X, T, Y, nU, nOuter, nMHInner, nESInner)`
 * (src/inference.jl:4-379, called from samplePosterior, src/driver.jl:59-69) for n_chains chains with host buffers:
 * samples_out [nOuter][n_chains][stride]; accepts/ess_evals may be NULL. */
int gpslc_posterior(gpslc_ctx* ctx, const gpslc_data* data, const gpslc_prior* prior, const gpslc_opts* opts,
                    double* samples_out, unsigned long long* accepts, unsigned long long* ess_evals);

/* ---- L4: posterior-predictive treatment effects -------------------------------------------------------------- */

/* ITEDistributions + ITEsamples (src/estimation.jl:66-109; conditionalITE :36-50; likelihoodDistribution,
 * src/likelihood.jl:8-174) for every (doT, chain, retained sample). `samples` is the sampler's packed buffer
 * [n_outer][n_chains][stride]; ret_idx holds the R retained 0-based outer indices (the reference's
 * nBurnIn:stepSize:nOuter is 1-based and includes nBurnIn, src/estimation.jl:72,78). jitter = predictionCovarianceNoise.
 * Only data->{n,nX,nU,X,T,Y} are read (X is always the observed data, src/estimation.jl:59).
 * Outputs (each may be NULL):
 *   meanITE [n_doT][n_chains][R][n]
 *   covITE  [n_doT][n_chains][R][n][n]      Symmetric(CovITE) + jitter*I   (src/estimation.jl:82)
 *   ite     [n_doT][n_chains][R*spp][n]     column j = r*spp + s is one draw N(MeanITE_r, CovITE_r) — for one chain this
 *                                           is the column-major n x (R*spp) matrix sampleITE returns (src/driver.jl:86-89)
 *   info    [n_doT][n_chains][R]
 * One Cholesky of the 2n x 2n augmented matrix per (doT, chain, sample); 8 n^3 / 3 flops. */
int gpslc_ite(gpslc_ctx* ctx, int loc, const gpslc_data* data, const double* samples, int n_outer, int n_chains, int stride,
              const int* ret_idx, int R, const double* doT, int n_doT, double jitter, int spp, uint64_t seed, int chain_offset,
              double* meanITE, double* covITE, double* ite, int* info);

/* SATEDistributions + SATEsamples (src/estimation.jl:116-163) without forming CovITE: one n x n Cholesky with two
 * right-hand sides per (doT, chain, sample). var_as_std=1 reproduces the reference's `normal(mean, var)` call, which
 * uses the variance as a standard deviation; 0 draws with sqrt(var).
 * Outputs: meanSATE, varSATE [n_doT][n_chains][R]; sate [n_doT][n_chains][R*spp] (for one chain: what sampleSATE
 * returns, src/driver.jl:108-111); info [n_doT][n_chains][R]. */
int gpslc_sate(gpslc_ctx* ctx, int loc, const gpslc_data* data, const double* samples, int n_outer, int n_chains, int stride,
               const int* ret_idx, int R, const double* doT, int n_doT, double jitter, int spp, uint64_t seed, int chain_offset,
               int var_as_std, double* meanSATE, double* varSATE, double* sate, int* info);

/* gpslc_ite / gpslc_sate for a SLICE of a counterfactual sweep (predictCounterfactualEffects, src/prediction.jl:23-36, sharded
 * over GPUs — BASELINE config c5: 256 doT values, 32 per GPU): doT[0..n_doT) are elements dot_offset.. of the full sweep. The
 * draws' Philox streams are keyed by the global doT index, so the slices concatenate to exactly what the unsharded call
 * returns. All other arguments and outputs as in gpslc_ite / gpslc_sate. */
int gpslc_ite_slice(gpslc_ctx* ctx, int loc, const gpslc_data* data, const double* samples, int n_outer, int n_chains, int stride,
                    const int* ret_idx, int R, const double* doT, int n_doT, int dot_offset, double jitter, int spp, uint64_t seed,
                    int chain_offset, double* meanITE, double* covITE, double* ite, int* info);
int gpslc_sate_slice(gpslc_ctx* ctx, int loc, const gpslc_data* data, const double* samples, int n_outer, int n_chains, int stride,
                     const int* ret_idx, int R, const double* doT, int n_doT, int dot_offset, double jitter, int spp, uint64_t seed,
                     int chain_offset, int var_as_std, double* meanSATE, double* varSATE, double* sate, int* info);

/* predictCounterfactualEffects (src/prediction.jl:23-36) + summarizeEstimates (src/driver.jl:129-149) in one call: the draws of
 * gpslc_ite_slice stay in the library's device arena and only the per-individual statistics come back —
 *   summary [n_doT][n_chains][n][3] = Mean, LowerBound, UpperBound of the R*spp draws of each (doT, chain)
 * (BASELINE config c5: 256 doT x 8192 individuals x 10 draws = 168 MB of draws never cross PCIe). */
int gpslc_ite_summary(gpslc_ctx* ctx, int loc, const gpslc_data* data, const double* samples, int n_outer, int n_chains, int stride,
                      const int* ret_idx, int R, const double* doT, int n_doT, int dot_offset, double jitter, int spp, uint64_t seed,
                      int chain_offset, double credible_interval, double* summary, int* info);

/* The subgroup effect curve of the reference's documented workflow (docs/src/index.md:101-114):
 *   ite, doT = predictCounterfactualEffects(g, nSamples); sate = mean(ite[:, idx, :], dims=2)[:, 1, :]; summarizeEstimates(sate)
 * in one call; the draws and the subgroup averages are produced and reduced in HBM.
 *   mask    [n] bytes (host memory in either mode), non-zero = the individual belongs to the subgroup; at least one
 *   sate    [n_chains][R*spp][n_doT]  subgroup average of every draw — per chain the column-major nDoT x nSamples matrix the
 *                                     example hands to summarizeEstimates (may be NULL)
 *   summary [n_chains][n_doT][3]      Mean, LowerBound, UpperBound over the R*spp draws, one row per doT value
 * doT / dot_offset / info as in gpslc_ite_slice (a shard covers a contiguous block of doT values). */
int gpslc_ite_subset_summary(gpslc_ctx* ctx, int loc, const gpslc_data* data, const double* samples, int n_outer, int n_chains,
                             int stride, const int* ret_idx, int R, const double* doT, int n_doT, int dot_offset, double jitter,
                             int spp, uint64_t seed, int chain_offset, const unsigned char* mask, double credible_interval,
                             double* sate, double* summary, int* info);
/* mean(samples[:, idx, :], dims=2) alone (docs/src/index.md:104-108): samples [batch][m][n] (the layout gpslc_ite writes) ->
 * out [batch][m]; HBM-bound, every sample read once. */
int gpslc_subset_mean(gpslc_ctx* ctx, int loc, const double* samples, int batch, int m, int n, const unsigned char* mask,
                      double* out);

/* summarizeEstimates (src/driver.jl:129-149): for every individual the mean and the (1-ci)/2 and 1-(1-ci)/2 quantiles of its
 * m samples, Julia's default `quantile` (linear interpolation between order statistics, type 7; KAT test/driver.jl:54-71).
 *   samples [batch][m][n]  — the layout gpslc_ite writes (`ite` for one doT and one chain is one batch element with
 *                            m = R*spp; for one doT and all chains pooled, m = n_chains*R*spp)
 *   out     [batch][n][3]  — Mean, LowerBound, UpperBound
 * With loc = GPSLC_DEVICE a counterfactual sweep is summarised where gpslc_ite left it in HBM (BASELINE config c5 never
 * ships its 168 MB of draws to the host). Any m: up to 8192 samples per individual are sorted in shared memory, more (e.g. the draws
 * of all chains pooled) go through a radix selection of the four order statistics. */
int gpslc_summarize(gpslc_ctx* ctx, int loc, const double* samples, int batch, int m, int n, double credible_interval,
                    double* out);

#ifdef __cplusplus
}
#endif
#endif /* GPSLC_H */
