// Matrix-entry generators for factor_run (see factor.cuh): the covariance build of the reference
// (src/kernel.jl:13-59: rbfKernelLog summed over feature groups, then processCov) evaluated directly in the
// accumulator layout of the Cholesky tiles, so K never exists in HBM on the sampler path.
#pragma once
#include "common.cuh"
#include "factor.cuh"

#ifndef GPSLC_STAGE_COLS
#define GPSLC_STAGE_COLS 1
#endif

namespace gpslc {

// exp(-a) for a >= 0 without branches, so that the 16 independent evaluations of a strip interleave in the FP64 pipe instead
// of running as 16 serial dependent chains through the library routine's special-case branches. Cody-Waite reduction
// x = k ln2 + r, |r| <= ln2/2, degree-13 Taylor polynomial in Horner form (truncation 4e-18), result scaled through the
// exponent field; values below 2^-1020 are flushed to zero (the library would return denormals). Max error ~1.5 ulp.
__device__ __forceinline__ double exp_neg(double a) {
    const double x = -a;
    const double kd = rint(x * 1.4426950408889634074);
    double r = fma(kd, -6.93147180369123816490e-01, x);
    r = fma(kd, -1.90821492927058770002e-10, r);
    double p = 1.6059043836821613e-10;
    p = fma(p, r, 2.08767569878681e-09);
    p = fma(p, r, 2.505210838544172e-08);
    p = fma(p, r, 2.755731922398589e-07);
    p = fma(p, r, 2.7557319223985893e-06);
    p = fma(p, r, 2.48015873015873e-05);
    p = fma(p, r, 1.984126984126984e-04);
    p = fma(p, r, 1.388888888888889e-03);
    p = fma(p, r, 8.333333333333333e-03);
    p = fma(p, r, 4.1666666666666664e-02);
    p = fma(p, r, 1.6666666666666666e-01);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const int k = (int)kd;
    const double v = __hiloint2double(__double2hiint(p) + (k << 20), __double2loint(p));
    return (k < -1020) ? 0.0 : v;
}

// exp(-a), a >= 0, with fewer FP64-pipe instructions (11 instead of ~20: the covariance generation shares that pipe with the
// DMMAs): x = -a = (32 e + i) ln2/32 + r with |r| <= ln2/64, exp(x) = 2^e * 2^(i/32) * P6(r). 2^(i/32) comes from a 32-entry
// table in shared memory (correctly rounded entries), the integer part of x 32/ln2 from the low word of the magic-number sum,
// the Taylor polynomial of degree 6 truncates at 3.5e-18. Max error ~2 ulp. Results below 2^-1020 are flushed to zero.
__device__ __forceinline__ double exp_neg_tab(double a, const double* tab) {
    const double MAGIC = 6755399441055744.0;   // 1.5 * 2^52: the sum's low word is round(x * 32/ln2)
    const double t = fma(-a, 46.16624130844683, MAGIC);
    const int k = __double2loint(t);
    const double kd = t - MAGIC;
    double r = fma(kd, -0.02166084938653512, -a);          // ln2/32, upper 32 bits (kd * hi is exact)
    r = fma(kd, -5.9631716539705866e-12, r);
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    const double v = tab[k & 31] * p;
    const int e = k >> 5;
    const double out = __hiloint2double(__double2hiint(v) + (e << 20), __double2loint(v));
    // a > 745 (integer compare on the high word: a >= 0) would wrap k; anything below 2^-1020 is flushed
    return (__double2hiint(a) >= 0x40874800 || e < -1020) ? 0.0 : out;
}

// K[r][c] = scale * exp(-sum_d w_d (f_d[r]-f_d[c])^2) + noise*[r==c],  w_d = 1/ls_d^2  (no 1/2, lengthscale squared:
// src/kernel.jl:17). Rows/cols >= n are identity padding.
struct RbfSpec {
    const double* feat[DMAX];
    double w[DMAX];
    double sw[DMAX];    // sqrt(w_d) = 1 / lengthscale_d: the tile generator works on pre-scaled features, (sw f_r - sw f_c)^2
    const double* y[MAXRHS];
    double scale, noise;
    int D, n;
};

// The strip() fast path and the element-wise one() (ragged edges) must give bit-identical values for the same entry: which of
// the two evaluates an entry depends on how the row blocks are tiled (one CTA or a cluster team, 64- or 128-row tiles), and a
// chain is required not to depend on that (nor, through the team size, on how many chains run beside it). Both therefore use
// the same operations in the same order, written with explicit round-to-nearest intrinsics so that the compiler cannot contract
// a multiply-add in one path and not in the other.
struct RbfGen {
    const RbfSpec* s;
    const double* tab;          // FactorSmem::exp2tab
    // element-wise path (ragged edges only): deliberately NOT inlined - sixteen inlined copies per strip instance were half of the
    // sampler kernel's 290 KB of code, and at small n the instruction fetch misses show up as `no_instruction` stalls
    __device__ __noinline__ double one(int r, int c) const {
        if (r >= s->n || c >= s->n) return (r == c) ? 1.0 : 0.0;
        double a = 0.0;
#pragma unroll 1
        for (int d = 0; d < s->D; d++) {
            const double* p = s->feat[d];
            const double w = s->sw[d];
            const double t = __dsub_rn(__dmul_rn(__ldg(p + r), w), __dmul_rn(__ldg(p + c), w));
            a = __fma_rn(t, t, a);
        }
        return __fma_rn(s->scale, exp_neg_tab(a, tab), (r == c) ? s->noise : 0.0);
    }
    __device__ __forceinline__ void quad(int r0, int r1, int c, double& v00, double& v01, double& v10, double& v11) const {
        v00 = one(r0, c); v01 = one(r0, c + 1); v10 = one(r1, c); v11 = one(r1, c + 1);
    }
    // NI column pairs (c0 + 8*ni, +1) for rows r0 and (unless ONE_ROW) r1: the accumulator layout of one warp tile.
    // Row features are loaded once per dimension for all NI pairs; all loads go through the read-only path.
    // cooperative: feature values of the panel's 64 columns into shared memory (zero beyond n)
    __device__ __forceinline__ void stage_cols(int col0, double* cf) const {
        const int D = s->D;
        if (!GPSLC_STAGE_COLS || D > CF_DIMS) return;
        for (int i = threadIdx.x; i < D * NB; i += blockDim.x) {
            const int d = i >> 6, c = col0 + (i & 63);
            cf[i] = (c < s->n) ? __dmul_rn(__ldg(s->feat[d] + c), s->sw[d]) : 0.0;    // pre-scaled by 1 / lengthscale
        }
    }
    template <int NI, bool ONE_ROW>
    __device__ __forceinline__ void strip(int r0, int r1, int c0, double (&v)[2][NI][2], const double* cf, int cl) const {
        const int n = s->n;
        const int D = s->D;
        // fast path: all entries inside the matrix and the panel's column features staged in shared memory (D <= CF_DIMS); otherwise
        // (ragged edge, or more dimensions than the staging area holds) the element-wise path below, which gives the same values
        // (a ONE_ROW caller's r1 is a dummy: testing it sent the warp that owns the last 8 rows of the matrix down the element-wise path)
        if ((ONE_ROW || r1 < n) && r0 < n && c0 + 8 * (NI - 1) + 1 < n && GPSLC_STAGE_COLS && D <= CF_DIMS) {
            double a[2][NI][2];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) { a[0][ni][0] = 0.0; a[0][ni][1] = 0.0; a[1][ni][0] = 0.0; a[1][ni][1] = 0.0; }
#pragma unroll 2
            for (int d = 0; d < D; d++) {
                const double w = s->sw[d];
                const double z0 = __dmul_rn(__ldg(s->feat[d] + r0), w);
                const double z1 = ONE_ROW ? z0 : __dmul_rn(__ldg(s->feat[d] + r1), w);
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
            const double sc = s->scale, nz = s->noise;
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                const int c = c0 + 8 * ni;
                v[0][ni][0] = __fma_rn(sc, exp_neg_tab(a[0][ni][0], tab), (r0 == c) ? nz : 0.0);
                v[0][ni][1] = __fma_rn(sc, exp_neg_tab(a[0][ni][1], tab), (r0 == c + 1) ? nz : 0.0);
                if (!ONE_ROW) {
                    v[1][ni][0] = __fma_rn(sc, exp_neg_tab(a[1][ni][0], tab), (r1 == c) ? nz : 0.0);
                    v[1][ni][1] = __fma_rn(sc, exp_neg_tab(a[1][ni][1], tab), (r1 == c + 1) ? nz : 0.0);
                }
            }
        } else {
            // slow path, kept small: a rolled loop over the entries through a local array (dynamic indexing is fine here)
            double tmp[2 * NI * 2];
#pragma unroll 1
            for (int idx = 0; idx < (ONE_ROW ? 2 : 4) * NI; idx++) {
                const int rr = idx / (2 * NI), ni = (idx >> 1) % NI, e = idx & 1;
                tmp[idx] = one(rr ? r1 : r0, c0 + 8 * ni + e);
            }
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                v[0][ni][0] = tmp[2 * ni]; v[0][ni][1] = tmp[2 * ni + 1];
                if (!ONE_ROW) { v[1][ni][0] = tmp[2 * NI + 2 * ni]; v[1][ni][1] = tmp[2 * NI + 2 * ni + 1]; }
            }
        }
    }
    __device__ __forceinline__ double rhs(int which, int r) const { return (r < s->n) ? s->y[which][r] : 0.0; }
};

// strip() for generators that only provide quad()
#define GPSLC_GENERIC_STRIP                                                                                         \
    template <int NI, bool ONE_ROW>                                                                                 \
    __device__ __forceinline__ void strip(int r0, int r1, int c0, double (&v)[2][NI][2], const double*, int) const { \
        _Pragma("unroll") for (int ni = 0; ni < NI; ni++)                                                           \
            quad(r0, r1, c0 + 8 * ni, v[0][ni][0], v[0][ni][1], v[1][ni][0], v[1][ni][1]);                          \
    }                                                                                                               \
    __device__ __forceinline__ void stage_cols(int, double*) const {}

// Dense symmetric matrix read from memory (column-major, lower triangle referenced), for the standalone
// gpslc_chol_logpdf primitive.
struct DenseGen {
    const double* K; const double* y[MAXRHS]; int n; int ld;
    __device__ __forceinline__ double one(int r, int c) const {
        if (r >= n || c >= n) return (r == c) ? 1.0 : 0.0;
        return (r >= c) ? K[(size_t)c * ld + r] : K[(size_t)r * ld + c];
    }
    __device__ __forceinline__ void quad(int r0, int r1, int c, double& v00, double& v01, double& v10, double& v11) const {
        v00 = one(r0, c); v01 = one(r0, c + 1); v10 = one(r1, c); v11 = one(r1, c + 1);
    }
    __device__ __forceinline__ double rhs(int which, int r) const { return (r < n) ? y[which][r] : 0.0; }
    GPSLC_GENERIC_STRIP
};

}  // namespace gpslc
