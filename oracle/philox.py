"""Counter-based RNG shared (as a specification) by the oracle and the CUDA path.

The reference draws from Julia's global Xoshiro256++, which a many-chain GPU sampler cannot
reproduce; parity with the reference is therefore distributional. Between THIS oracle and the CUDA kernels,
however, the random streams are identical by construction: Philox4x32-10 (Salmon et al., SC'11) keyed by the
user seed, with the 128-bit counter naming (block, a, chain, b) — see ``stream_b``. The transformations
(uniform, Box-Muller normal, Marsaglia-Tsang gamma) are restated here exactly as csrc/rng.cuh implements them.

Test infrastructure only (see oracle/__init__.py).
"""
import math

import numpy as np

M0 = 0xD2511F53
M1 = 0xCD9E8D57
W0 = 0x9E3779B9
W1 = 0xBB67AE85
MASK = 0xFFFFFFFF

# purpose tags (top 4 bits of counter word 3)
TAG_INIT_PARAM = 1
TAG_INIT_VEC = 2
TAG_MH_PROP = 3
TAG_MH_ACC = 4
TAG_ESS_NU = 5
TAG_ESS_SCALAR = 6
TAG_ITE = 7
TAG_SATE = 8
TAG_INIT_XMODEL = 9


def stream_b(tag, it):
    """Counter word 3: purpose tag in the top 4 bits, iteration index below."""
    assert 0 <= it < (1 << 28)
    return ((tag & 0xF) << 28) | it


def philox4x32(c0, c1, c2, c3, k0, k1):
    """One Philox4x32-10 block. All arguments Python ints < 2**32. Returns 4 ints."""
    for _ in range(10):
        p0 = M0 * c0
        p1 = M1 * c2
        hi0, lo0 = p0 >> 32, p0 & MASK
        hi1, lo1 = p1 >> 32, p1 & MASK
        c0, c1, c2, c3 = (hi1 ^ c1 ^ k0) & MASK, lo1, (hi0 ^ c3 ^ k1) & MASK, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def philox4x32_vec(c0, c1, c2, c3, k0, k1):
    """Vectorised over c0 (uint64 array holding 32-bit values); other words scalar ints."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.full_like(c0, c1)
    c2 = np.full_like(c0, c2)
    c3 = np.full_like(c0, c3)
    m = np.uint64(MASK)
    s32 = np.uint64(32)
    for _ in range(10):
        p0 = np.uint64(M0) * c0
        p1 = np.uint64(M1) * c2
        hi0, lo0 = p0 >> s32, p0 & m
        hi1, lo1 = p1 >> s32, p1 & m
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0)) & m, lo1, (hi0 ^ c3 ^ np.uint64(k1)) & m, lo0
        k0 = (k0 + W0) & MASK
        k1 = (k1 + W1) & MASK
    return c0, c1, c2, c3


def u01(lo, hi):
    """53-bit uniform in (0,1): ((hi:lo) >> 11 + 0.5) * 2^-53."""
    x = ((hi << 32) | lo) >> 11
    return (x + 0.5) * (1.0 / 9007199254740992.0)


class Stream:
    """A named Philox stream: fixed (a, chain, b) and key=seed, block counter advancing from 0."""

    def __init__(self, seed, chain, a, b):
        self.k0 = seed & MASK
        self.k1 = (seed >> 32) & MASK
        self.chain = chain & MASK
        self.a = a & MASK
        self.b = b & MASK
        self.block = 0

    def next_block(self):
        w = philox4x32(self.block & MASK, self.a, self.chain, self.b, self.k0, self.k1)
        self.block += 1
        return w

    def uniform_pair(self):
        w = self.next_block()
        return u01(w[0], w[1]), u01(w[2], w[3])

    def uniform(self):
        return self.uniform_pair()[0]

    def normal(self):
        """One normal per block (Box-Muller cosine branch)."""
        u1, u2 = self.uniform_pair()
        return math.sqrt(-2.0 * math.log(u1)) * math.cos(2.0 * math.pi * u2)

    def gamma(self, shape):
        """Marsaglia-Tsang (2000); one normal block + one uniform block per attempt; shape < 1 through the boost
        Gamma(a) = Gamma(a+1) * U^(1/a). Same bounded loop as csrc/rng.cuh."""
        assert shape > 0.0
        a = shape + 1.0 if shape < 1.0 else shape
        d = a - 1.0 / 3.0
        c = 1.0 / math.sqrt(9.0 * d)
        g = d
        for _ in range(256):
            x = self.normal()
            u = self.uniform()
            v = 1.0 + c * x
            if v <= 0.0:
                continue
            v = v * v * v
            g = d * v
            if math.log(u) < 0.5 * x * x + d - d * v + d * math.log(v):
                break
        if shape < 1.0:
            g *= self.uniform() ** (1.0 / shape)
        return g

    def inv_gamma(self, shape, scale):
        """Gen `inv_gamma(shape, scale)` == scale / Gamma(shape, 1)."""
        return scale / self.gamma(shape)

    def normal_vector(self, n):
        """n normals: element i uses block i//2, cosine branch for even i, sine branch for odd i."""
        nb = (n + 1) // 2
        blocks = np.arange(self.block, self.block + nb, dtype=np.uint64)
        w0, w1, w2, w3 = philox4x32_vec(blocks, self.a, self.chain, self.b, self.k0, self.k1)
        self.block += nb
        x1 = ((w1 << np.uint64(32)) | w0) >> np.uint64(11)
        x2 = ((w3 << np.uint64(32)) | w2) >> np.uint64(11)
        u1 = (x1.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
        u2 = (x2.astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)
        r = np.sqrt(-2.0 * np.log(u1))
        z = np.empty(2 * nb)
        z[0::2] = r * np.cos(2.0 * np.pi * u2)
        z[1::2] = r * np.sin(2.0 * np.pi * u2)
        return z[:n]
