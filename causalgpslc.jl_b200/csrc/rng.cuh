// Counter-based RNG: Philox4x32-10 keyed by the user seed; the specification is shared with oracle/philox.py
// (same counter naming, same uniform/normal/gamma transformations) so that the CPU oracle and the CUDA chains
// consume identical random streams. Replaces the reference's use of Julia's global Xoshiro256++.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace gpslc {

enum RngTag : uint32_t {
    TAG_INIT_PARAM = 1, TAG_INIT_VEC = 2, TAG_MH_PROP = 3, TAG_MH_ACC = 4, TAG_ESS_NU = 5, TAG_ESS_SCALAR = 6,
    TAG_ITE = 7, TAG_SATE = 8, TAG_INIT_XMODEL = 9
};

__host__ __device__ inline uint32_t stream_b(uint32_t tag, uint32_t it) { return ((tag & 0xFu) << 28) | (it & 0x0FFFFFFFu); }

struct Philox4 { uint32_t x, y, z, w; };

__host__ __device__ inline Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
    const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)M0 * c0, p1 = (uint64_t)M1 * c2;
        uint32_t hi0 = (uint32_t)(p0 >> 32), lo0 = (uint32_t)p0, hi1 = (uint32_t)(p1 >> 32), lo1 = (uint32_t)p1;
        uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += W0; k1 += W1;
    }
    return Philox4{c0, c1, c2, c3};
}

__host__ __device__ inline double u01(uint32_t lo, uint32_t hi) {
    uint64_t x = (((uint64_t)hi << 32) | lo) >> 11;
    return ((double)x + 0.5) * (1.0 / 9007199254740992.0);
}

// A named stream: (a, chain, b) fixed, block counter advancing from 0.
struct Stream {
    uint32_t k0, k1, chain, a, b, block;
    __host__ __device__ Stream(uint64_t seed, uint32_t chain_, uint32_t a_, uint32_t b_)
        : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)), chain(chain_), a(a_), b(b_), block(0) {}
    __host__ __device__ Philox4 next() { return philox4x32_10(block++, a, chain, b, k0, k1); }
    __host__ __device__ Philox4 at(uint32_t blk) const { return philox4x32_10(blk, a, chain, b, k0, k1); }
    __host__ __device__ void uniform_pair(double& u1, double& u2) { Philox4 p = next(); u1 = u01(p.x, p.y); u2 = u01(p.z, p.w); }
    __host__ __device__ double uniform() { double u1, u2; uniform_pair(u1, u2); return u1; }
    __host__ __device__ double normal() {
        double u1, u2; uniform_pair(u1, u2);
        return sqrt(-2.0 * log(u1)) * cos(2.0 * 3.14159265358979323846 * u2);
    }
    // Marsaglia-Tsang (2000): one normal block + one uniform block per attempt. Shapes below 1 use the boost
    // Gamma(a) = Gamma(a + 1) * U^(1/a) (one more uniform block, drawn after the accepted attempt). The rejection loop accepts with
    // probability > 0.95 per attempt; it is bounded so that a NaN shape can never hang a kernel (the last candidate is returned).
    __host__ __device__ double gamma(double shape) {
        const double a = (shape < 1.0) ? shape + 1.0 : shape;
        const double d = a - 1.0 / 3.0, c = 1.0 / sqrt(9.0 * d);
        double g = d;
        for (int attempt = 0; attempt < 256; attempt++) {
            double x = normal();
            double u = uniform();
            double v = 1.0 + c * x;
            if (v <= 0.0) continue;
            v = v * v * v;
            g = d * v;
            if (log(u) < 0.5 * x * x + d - d * v + d * log(v)) break;
        }
        if (shape < 1.0) g *= pow(uniform(), 1.0 / shape);
        return g;
    }
    __host__ __device__ double inv_gamma(double shape, double scale) { return scale / gamma(shape); }
    // element i of the stream's normal vector: block i/2, cosine branch for even i, sine branch for odd i
    __host__ __device__ double normal_at(uint32_t i) const {
        Philox4 p = at(i >> 1);
        double u1 = u01(p.x, p.y), u2 = u01(p.z, p.w);
        double r = sqrt(-2.0 * log(u1)), ang = 2.0 * 3.14159265358979323846 * u2;
        return (i & 1) ? r * sin(ang) : r * cos(ang);
    }
};

}  // namespace gpslc
