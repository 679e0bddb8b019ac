// Matrix-entry generators for factor_run (see factor.cuh): the covariance build of the reference
// (src/kernel.jl:13-59: rbfKernelLog summed over feature groups, then processCov) evaluated directly in the
// accumulator layout of the Cholesky tiles, so K never exists in HBM on the sampler path.
#pragma once
#include "common.cuh"
#include "factor.cuh"

namespace gpslc {

// K[r][c] = scale * exp(-sum_d w_d (f_d[r]-f_d[c])^2) + noise*[r==c],  w_d = 1/ls_d^2  (no 1/2, lengthscale squared:
// src/kernel.jl:17). Rows/cols >= n are identity padding.
struct RbfSpec {
    const double* feat[DMAX];
    double w[DMAX];
    const double* y[MAXRHS];
    double scale, noise;
    int D, n;
};

struct RbfGen {
    const RbfSpec* s;
    __device__ __forceinline__ double one(int r, int c) const {
        if (r >= s->n || c >= s->n) return (r == c) ? 1.0 : 0.0;
        double a = 0.0;
        for (int d = 0; d < s->D; d++) {
            const double* p = s->feat[d];
            const double t = p[r] - p[c];
            a = fma(t * s->w[d], t, a);
        }
        double v = s->scale * exp(-a);
        if (r == c) v += s->noise;
        return v;
    }
    __device__ __forceinline__ void quad(int r0, int r1, int c, double& v00, double& v01, double& v10, double& v11) const {
        const int n = s->n;
        if (r1 < n && r0 < n && c + 1 < n) {
            double a00 = 0.0, a01 = 0.0, a10 = 0.0, a11 = 0.0;
            const int D = s->D;
            for (int d = 0; d < D; d++) {
                const double* p = s->feat[d];
                const double w = s->w[d];
                const double z0 = p[r0], z1 = p[r1], c0 = p[c], c1 = p[c + 1];
                double t;
                t = z0 - c0; a00 = fma(t * w, t, a00);
                t = z0 - c1; a01 = fma(t * w, t, a01);
                t = z1 - c0; a10 = fma(t * w, t, a10);
                t = z1 - c1; a11 = fma(t * w, t, a11);
            }
            const double sc = s->scale;
            v00 = sc * exp(-a00); v01 = sc * exp(-a01); v10 = sc * exp(-a10); v11 = sc * exp(-a11);
            const double nz = s->noise;
            if (r0 == c) v00 += nz;
            if (r0 == c + 1) v01 += nz;
            if (r1 == c) v10 += nz;
            if (r1 == c + 1) v11 += nz;
        } else {
            v00 = one(r0, c); v01 = one(r0, c + 1); v10 = one(r1, c); v11 = one(r1, c + 1);
        }
    }
    // NI column pairs (c0 + 8*ni, +1) for rows r0 and (unless ONE_ROW) r1: the accumulator layout of one warp tile.
    // Row features are loaded once per dimension for all NI pairs; all loads go through the read-only path.
    template <int NI, bool ONE_ROW>
    __device__ __forceinline__ void strip(int r0, int r1, int c0, double (&v)[2][NI][2]) const {
        const int n = s->n;
        if (r1 < n && r0 < n && c0 + 8 * (NI - 1) + 1 < n) {
            double a[2][NI][2];
#pragma unroll
            for (int ni = 0; ni < NI; ni++) { a[0][ni][0] = 0.0; a[0][ni][1] = 0.0; a[1][ni][0] = 0.0; a[1][ni][1] = 0.0; }
            const int D = s->D;
#pragma unroll 2
            for (int d = 0; d < D; d++) {
                const double* p = s->feat[d];
                const double w = s->w[d];
                const double z0 = __ldg(p + r0);
                const double z1 = ONE_ROW ? z0 : __ldg(p + r1);
#pragma unroll
                for (int ni = 0; ni < NI; ni++) {
                    const double c0v = __ldg(p + c0 + 8 * ni), c1v = __ldg(p + c0 + 8 * ni + 1);
                    double t;
                    t = z0 - c0v; a[0][ni][0] = fma(t * w, t, a[0][ni][0]);
                    t = z0 - c1v; a[0][ni][1] = fma(t * w, t, a[0][ni][1]);
                    if (!ONE_ROW) {
                        t = z1 - c0v; a[1][ni][0] = fma(t * w, t, a[1][ni][0]);
                        t = z1 - c1v; a[1][ni][1] = fma(t * w, t, a[1][ni][1]);
                    }
                }
            }
            const double sc = s->scale, nz = s->noise;
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                const int c = c0 + 8 * ni;
                v[0][ni][0] = sc * exp(-a[0][ni][0]) + ((r0 == c) ? nz : 0.0);
                v[0][ni][1] = sc * exp(-a[0][ni][1]) + ((r0 == c + 1) ? nz : 0.0);
                if (!ONE_ROW) {
                    v[1][ni][0] = sc * exp(-a[1][ni][0]) + ((r1 == c) ? nz : 0.0);
                    v[1][ni][1] = sc * exp(-a[1][ni][1]) + ((r1 == c + 1) ? nz : 0.0);
                }
            }
        } else {
#pragma unroll
            for (int ni = 0; ni < NI; ni++) {
                const int c = c0 + 8 * ni;
                v[0][ni][0] = one(r0, c); v[0][ni][1] = one(r0, c + 1);
                if (!ONE_ROW) { v[1][ni][0] = one(r1, c); v[1][ni][1] = one(r1, c + 1); }
            }
        }
    }
    __device__ __forceinline__ double rhs(int which, int r) const { return (r < s->n) ? s->y[which][r] : 0.0; }
};

// strip() for generators that only provide quad()
#define GPSLC_GENERIC_STRIP                                                                                         \
    template <int NI, bool ONE_ROW>                                                                                 \
    __device__ __forceinline__ void strip(int r0, int r1, int c0, double (&v)[2][NI][2]) const {                    \
        _Pragma("unroll") for (int ni = 0; ni < NI; ni++)                                                           \
            quad(r0, r1, c0 + 8 * ni, v[0][ni][0], v[0][ni][1], v[1][ni][0], v[1][ni][1]);                          \
    }

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
