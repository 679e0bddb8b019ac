// One-CTA blocked Cholesky + forward solves for one SPD matrix whose entries are produced on the fly by a
// generator functor (RBF covariance build fused into the factorisation, or a dense matrix read, or the augmented
// ITE matrix). Replaces, for one (chain, factor) pair, the reference's
//   rbfKernelLog + processCov            (src/kernel.jl:24-59)
//   Distributions.logpdf(MvNormal(0,K),y) (reached through Gen `mvnormal` at src/model_likelihood.jl:30-118)
// i.e. LAPACK dpotrf + dtrsv + log-diag sum.
//
// Algorithm: left-looking, 64-wide panels. For panel j
//   diag tile : C_jj = K_jj - sum_{J<j} L_jJ L_jJ^T          (DMMA m8n8k4, operands streamed by TMA bulk copies)
//               w_j  = y_j  - sum_{J<j} L_jJ z_J             (fused in the same k-loop, plain DFMA)
//   P2        : [L_jj ; z_j^T] = Cholesky of the bordered block [C_jj w_j ; w_j^T .] in shared memory, right-looking over
//               8-column tiles: 8x8 pivot tiles by one warp in registers, everything else 8x8x4 DMMAs; leaves the inverses of
//               the 8x8 diagonal tiles and -L_jj (below them) as B-fragment atoms for the row tiles
//   row tiles : C_Ij = K_Ij - sum_J L_IJ L_jJ^T  (128x64 per tile, non-inlined spill-free k-loop) ; L_Ij = C_Ij L_jj^-T by block
//               forward substitution over 8-column tiles (DMMA, A-fragments rebuilt from accumulator-layout values with quad
//               shuffles) ; L_Ij stored to the CTA's global scratch.
// L lives in global scratch in an "atom" layout: every 8(row) x 4(k) DMMA operand fragment is 256 contiguous bytes,
// fragments are grouped into 64(row) x 16(k) slabs of 8 KB so that one cp.async.bulk moves one pipeline operand.
// DESIGN.md §3.1 has the rationale of each piece and profiles/phase_timing_r01.md what it bought.
#pragma once
#include <type_traits>
#include "common.cuh"

namespace gpslc {

// Development aid (-DGPSLC_PHASE_TIMING, tools/gpu_phase_timing.py): cycles thread 0 of every CTA spends in each phase of
// factor_run, accumulated in a per-translation-unit device array. Compiles to nothing in the product build.
#ifdef GPSLC_PHASE_TIMING
static __device__ unsigned long long g_phase_cycles[64];   // 24..31 end-of-panel barrier wait per warp, 32..39 row k-loop time per warp,
                                                          // 40..47 operand wait per warp, 48..55 stage refills issued per warp
#define GP_PHASE_INIT() long long _pt = clock64()
#define GP_PHASE_MARK(k) do { if (threadIdx.x == 0) { const long long _n = clock64(); atomicAdd(&g_phase_cycles[k], (unsigned long long)(_n - _pt)); _pt = _n; } } while (0)
#else
#define GP_PHASE_INIT() do {} while (0)
#define GP_PHASE_MARK(k) do {} while (0)
#endif

constexpr int NB = 64;                 // panel width == row-block height
constexpr int KB = 16;                 // k extent of one pipeline slab
constexpr int K4S = KB / 4;            // 8x4 operand atoms along k per slab
constexpr int NSLAB = NB / KB;         // slabs per block
constexpr int SLAB_D = NB * KB;        // doubles per slab (8 KB)
constexpr int BLOCK_D = NB * NB;       // doubles per block (32 KB)
#ifndef GPSLC_CTAS
#define GPSLC_CTAS 2           // resident factor CTAs per SM the kernels are built for (3: experiment with a two-stage operand ring)
#endif
constexpr int CTAS_PER_SM = GPSLC_CTAS;
#ifndef GPSLC_STAGES
#define GPSLC_STAGES ((GPSLC_CTAS >= 3) ? 2 : 3)
#endif
constexpr int STAGES = GPSLC_STAGES;
constexpr int FWARPS = 8;
constexpr int FTHREADS = FWARPS * 32;
constexpr int STAGE_D = 3 * SLAB_D;    // A0 | A1 | B
constexpr int MAXRHS = 2;
// P2 workspace leading dimension (16-byte aligned rows). 68 = 4 mod 16 makes the A-fragment loads of P2 (LDS.64 at row g, column q)
// conflict-free per half-warp; with 66 every shared-memory access of P2's update was a 2-way conflict (ncu source page, c2: 3.5e9 of
// 20.6e9 wavefronts in excess). The 128-bit accumulator-layout accesses would want 8 mod 16 instead - no pitch serves both; measured
// at n = 256 / 128: 68 +1.5 % / +2.9 %, 72 +1.1 % / +2.1 %, c3 +0.15 %.
#ifndef GPSLC_CS_LD
#define GPSLC_CS_LD 68
#endif
constexpr int CS_LD = GPSLC_CS_LD;
constexpr int CS_ROWS = 72;            // 64 matrix rows + up to MAXRHS right-hand-side rows + zero padding to a full 8-row tile
constexpr int LINV_D = 72 * 32;        // atoms (n8, k4) with k4 <= 2*n8+1, row n8 starts at atom n8*(n8+1)
constexpr int CF_DIMS = (GPSLC_CTAS >= 3) ? 6 : 24;   // feature dimensions staged per panel (more dimensions fall back to global loads)

struct FactorOut {
    double logdet;      // log det K
    double gram[3];     // z0.z0, z0.z1, z1.z1 with z_i = L^-1 y_i
    int info;           // 0, or 1-based index of the first non-positive pivot (LAPACK convention)
};

struct __align__(128) FactorSmem {
    double stage[STAGES * STAGE_D];    // 72 KB; P2 aliases it as workspace while no copy is in flight
    double linv[LINV_D];               // 18 KB; what the row-tile epilogue needs of the current diagonal block, as the 72 8x4 B-fragment
                                       // atoms on/below the diagonal: inverses of the 8x8 diagonal tiles, minus L_jj below them
    double colfeat[CF_DIMS * NB];      // 12 KB; feature values of the current panel's 64 columns (covariance generation)
    double exp2tab[32];                // 2^(i/32) for the generators' exponential; must directly follow colfeat
    double p2buf[72];                  // pivot-warp broadcast buffer: an 8x8 column block of L and the 8 reciprocal pivots
    double wvec[MAXRHS][NB];           // w_j (pre-solve) / scratch
    double part[4][NB];                // per-row partial sums: log L_ii, z0.z0, z0.z1, z1.z1 (kept out of registers)
    double* snap; int snapJ, snap_n;   // snapshot hook parameters
    // shared-prefix mode (ITE: chol(Kp) computed once per posterior sample, see factor_run PRE): block rows < pre_split live in the
    // read-only scratch pre_lo, pre_linv holds the saved P2 outputs of their panels; save_linv: where a factorisation saves them
    const double* pre_lo; const double* pre_linv; double* save_linv; int pre_split;
    unsigned int* gbar;                // TEAM = 2: grid-wide barrier state {arrivals, generation}
    double red[32];
    unsigned long long full[STAGES];
    unsigned int freed[STAGES];        // warps that have finished with the slab in each stage (the last one refills it)
    FactorOut out;
    int info;
};

// ---------------------------------------------------------------------------------------------- PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
        : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {}
}
// TMA 1-D bulk copy global -> shared, completion signalled on an mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
// same copy with an L2 eviction-priority hint (createpolicy): the B slabs of a panel are re-read by every row tile of the panel,
// the A slabs are streamed once per panel
#ifndef GPSLC_L2_HINTS
#define GPSLC_L2_HINTS 1
#endif
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
    uint64_t p; asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p)); return p;
}
__device__ __forceinline__ void bulk_g2s_hint(void* dst, const void* src, uint32_t bytes, unsigned long long* bar, uint64_t policy) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst)), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }
// Stage release without an "empty" mbarrier and without a waiting producer: every warp counts itself out of a stage with an
// acq_rel shared-memory atomic; the warp that arrives last (and only that one) resets the counter and immediately issues the
// TMA copies that refill the stage, so no consumer warp ever blocks on the slowest one and the prefetch depth is the full ring.
__device__ __forceinline__ bool stage_release_is_last(unsigned int* cnt) {
    unsigned int old;
    asm volatile("atom.acq_rel.cta.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(smem_u32(cnt)) : "memory");
    if (old != FWARPS - 1) return false;
    *reinterpret_cast<volatile unsigned int*>(cnt) = 0;
    return true;
}

// Team mode (TEAM = 1): the CTAs of one thread-block cluster factor ONE matrix together (few large matrices, e.g. the
// n = 8192 counterfactual sweep, where one CTA per matrix would leave most SMs idle). The L scratch in global memory is shared by
// the team; the diagonal tile, P2 and the forward solve are computed redundantly by every CTA (each keeps its own Linv, z and
// partial sums), the row blocks below the diagonal are split across the team, and one cluster barrier (release/acquire) per
// panel publishes the new L blocks.
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_size() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_id_x() { uint32_t r; asm volatile("mov.u32 %0, %%clusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t cluster_count_x() { uint32_t r; asm volatile("mov.u32 %0, %%nclusterid.x;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_barrier() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TEAM = 2: every CTA of the (cooperatively launched, hence co-resident) grid works on ONE matrix - the single shared chol(Kp) of a
// large counterfactual sweep, where even the largest cluster would leave most SMs idle. Sense-reversing barrier on two words of global
// memory; the caller zeroes them before the launch.
__device__ __forceinline__ void grid_barrier(unsigned int* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int gen;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(gen) : "l"(bar + 1) : "memory");
        __threadfence();
        if (atomicAdd(bar, 1u) == gridDim.x - 1) {
            atomicExch(bar, 0u);
            __threadfence();
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(bar + 1), "r"(gen + 1) : "memory");
        } else {
            unsigned int cur;
            do { asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(bar + 1) : "memory"); } while (cur == gen);
        }
        __threadfence();
    }
    __syncthreads();
}
// end-of-phase synchronisation: generic-proxy writes (shared workspace, global L blocks) are ordered before later async-proxy
// (TMA bulk) reads, by this CTA or — in team mode — by any CTA of the cluster
template <int TEAM>
__device__ __forceinline__ void team_sync(unsigned int* gbar = nullptr) {
    fence_proxy_async();
    if constexpr (TEAM == 1) { cluster_barrier(); fence_proxy_async(); }
    else if constexpr (TEAM == 2) { grid_barrier(gbar); fence_proxy_async(); }
    else __syncthreads();
}

// FP64 tensor-core MMA (SASS: DMMA.8x8x4). A: lane holds A[lane/4][lane%4]; B: lane holds B[lane%4][lane/4];
// C/D: lane holds rows lane/4, cols 2*(lane%4)+{0,1}.
__device__ __forceinline__ void dmma(double (&d)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d[0]), "+d"(d[1]) : "d"(a), "d"(b));
}

// ---------------------------------------------------------------------------------------------- scratch layout
// blocks of row-block I are contiguous: (I,0), (I,1), ..., (I,I) — the k-loop of a row tile therefore streams one contiguous
// region per operand (slab t of row I lives at row_off(I) + t*SLAB_D)
__device__ __host__ inline size_t row_off(int I) { return ((size_t)I * (I + 1) / 2) * BLOCK_D; }
__device__ __host__ inline size_t block_off(int I, int J, int /*NRB*/) { return row_off(I) + (size_t)J * BLOCK_D; }
__host__ __device__ inline size_t scratch_doubles(int NRB, int /*NCB*/) { return row_off(NRB); }
// element (r,c) inside a block: [slab c/8][r/8][k4 (c%8)/4][ (r%8)*4 + c%4 ]
__device__ __host__ inline int elem_off(int r, int c) {
    return (c / KB) * SLAB_D + (r >> 3) * (K4S * 32) + ((c >> 2) % K4S) * 32 + (r & 7) * 4 + (c & 3);
}

// 2^(i/32), correctly rounded (see exp_neg_tab in gens.cuh)
static __device__ const double GPSLC_EXP2_TAB[32] = {
    1, 1.0218971486541166, 1.0442737824274138, 1.0671404006768237, 1.0905077326652577, 1.1143867425958924, 1.1387886347566916,
    1.1637248587775775, 1.189207115002721, 1.215247359980469, 1.241857812073484, 1.2690509571917332, 1.2968395546510096,
    1.3252366431597413, 1.3542555469368927, 1.383909881963832, 1.4142135623730951, 1.4451808069770467, 1.4768261459394993,
    1.5091644275934228, 1.5422108254079407, 1.5759808451078865, 1.6104903319492543, 1.6457554781539649, 1.681792830507429,
    1.7186192981224779, 1.7562521603732995, 1.7947090750031072, 1.8340080864093424, 1.8741676341103, 1.9152065613971474,
    1.9571441241754002};

__device__ inline void factor_smem_init(FactorSmem& sm) {
    if (threadIdx.x < 32) sm.exp2tab[threadIdx.x] = GPSLC_EXP2_TAB[threadIdx.x];
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; s++) { mbar_init(&sm.full[s], 1); sm.freed[s] = 0; }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------- P2
// Factor the 64x64 diagonal block held in Cs (row-major, CS_LD; the lower 8x8 tiles incl. full diagonal tiles are valid) and
// apply the same elimination to the right-hand-side rows 64..64+nrhs-1 of Cs (rows up to 71 are zero padding): afterwards
// Cs holds L_jj (lower triangle) and row 64+i holds z_i = L_jj^-1 w_i — the forward solve falls out of the factorisation
// of the bordered matrix. sm.linv receives what the row-tile epilogue needs, in B-fragment atoms: the inverses D_b of the
// eight 8x8 diagonal blocks of L_jj on the block diagonal, and MINUS the sub-diagonal part of L_jj elsewhere.
//
// Right-looking over 8-column blocks; per block: 8x8 Cholesky + inverse by warp 0 in registers (the only serial part),
// then X = A21 D^T and the trailing update A22 -= X X^T as 8x8x4 DMMAs spread over all warps. Everything that is not the
// 64-pivot chain is tensor work on purpose: while the sibling CTA streams DMMAs, the serial warp issues roughly one
// instruction per 7 cycles (measured), so this phase is paid for per instruction.
// Returns via sm.info the first non-positive pivot (1-based global column), if any.
// 1/sqrt(a) for a normal positive a in 5 FP64 instructions (the library routine takes 11): hardware seed (relative error
// < 2^-22) and one third-order step y (1 + e/2 + 3e^2/8), e = 1 - a y^2, which leaves ~2^-66 before rounding. Every FP64
// instruction of the pivot chain queues behind the sibling CTA's DMMAs, so the count is what matters here.
__device__ __forceinline__ double rsqrt_fast(double a) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
    const double t = a * y;
    const double e = fma(-t, y, 1.0);
    const double p = fma(0.375, e, 0.5);
    const double ye = y * e;
    return fma(ye, p, y);
}
__device__ __forceinline__ double neg_bits(double x) { return __hiloint2double(__double2hiint(x) ^ (int)0x80000000, __double2loint(x)); }

// GPSLC_PIVOT_RCP = 1: pivot recurrence through the reciprocal (4 dependent FP64 instructions per pivot instead of 7). Measured on
// B200, same box back to back: c3 1279.6 vs 1285.8 sweeps/s, c2 50.0 k vs 50.9 k - no gain (the pivot warp is not limited by the
// length of that chain but by issue slots behind the sibling CTA's DMMA stream), so it is off.
#ifndef GPSLC_PIVOT_RCP
#define GPSLC_PIVOT_RCP 0
#endif
__device__ inline void p2_factor_diag(FactorSmem& sm, double* Cs, int col0) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const unsigned FULL = 0xffffffffu;
    // warp index through a shuffle: tells the compiler the value is warp-uniform, so the shuffles inside `if (warp_u == 0)` are
    // plain SHFLs instead of WARPSYNC.COLLECTIVE sequences
    const int warp_u = __shfl_sync(FULL, warp, 0);
    GP_PHASE_INIT();
    // 8x8 Cholesky + inverse of diagonal tile kb by the calling warp (registers + shuffles)
    auto potf2 = [&](const int kb) {
        const int c0 = kb * 8;
        // lane r (= lane & 7) holds row r of the 8x8 block; afterwards column r of its inverse
        const int r = lane & 7;
        double a[8];
#ifdef GPSLC_PHASE_TIMING
        long long _q0 = clock64();
#define GP_SUB(k) do { if (threadIdx.x == 0) { const long long _n = clock64(); atomicAdd(&g_phase_cycles[k], (unsigned long long)(_n - _q0)); _q0 = _n; } } while (0)
#else
#define GP_SUB(k) do {} while (0)
#endif
#pragma unroll
        for (int k = 0; k < 8; k++) a[k] = (k <= r) ? Cs[(c0 + r) * CS_LD + c0 + k] : 0.0;
        GP_SUB(16);
        // Column k of L and 1/L_kk are broadcast through a 72-double shared-memory buffer (one 8-byte store per lane, then
        // 16-byte broadcast loads of column pairs) instead of one 2-instruction shuffle per element: the serial warp's time is
        // proportional to its instruction count (about one instruction per 6 cycles, measured), and shuffles were half of it.
        double* colbuf = sm.p2buf;          // [k][j] = L[j][k]
        double* rinvbuf = sm.p2buf + 64;    // [k] = 1 / L[k][k]
        double Lc[8][8];                    // register copy of the columns (j > k) for the inverse below
        // The pivot recurrence itself stays off both the shuffles and the buffer: every lane computes the next pivot,
        //   a_{k+1,k+1} <- a_{k+1,k+1} - (a_{k+1,k} / sqrt(a_kk))^2,
        // from two values of lane k+1 that are broadcast BEFORE 1/sqrt(a_kk) is known (the same expression, bit for bit, that lane
        // k+1 evaluates in its own update), so the serial chain per pivot is rsqrt -> mul -> fma.
        double akk = __shfl_sync(FULL, a[0], 0);
#pragma unroll
        for (int k = 0; k < 8; k++) {
            double d10 = 0.0, d11 = 0.0;
            if (k < 7) { d10 = __shfl_sync(FULL, a[k], k + 1); d11 = __shfl_sync(FULL, a[k + 1], k + 1); }
            // positive and finite? (integer test on the high word: no FP64-pipe instruction; NaN and denormal pivots fail it)
            const int hik = __double2hiint(akk);
            if (!(hik > 0 && hik < 0x7ff00000)) {
                if (lane == 0 && sm.info == 0) sm.info = col0 + c0 + k + 1;
                akk = 1.0;
            }
#if GPSLC_PIVOT_RCP
            // The chain from one pivot to the next goes through the RECIPROCAL, a_{k+1,k+1} - a_{k+1,k}^2 / a_kk: hardware seed and one
            // third-order step y (1 + e + e^2), e = 1 - a y (relative error seed^3 ~ 2^-60 before rounding), then one fma - four
            // dependent FP64 instructions instead of the seven of rsqrt -> multiply -> fma. 1/sqrt(a_kk), which scales the column, is
            // computed beside it and feeds only the column update, which has a pivot period of slack.
            double yr;
            asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(yr) : "d"(akk));
            const double er = fma(-akk, yr, 1.0);
            const double pr = fma(er, er, er);
            const double rcp = fma(yr, pr, yr);
            const double q10 = d10 * d10;
            const double ri = rsqrt_fast(akk);
            akk = fma(-q10, rcp, d11);          // next pivot
#else
            const double ri = rsqrt_fast(akk);
            const double l10 = d10 * ri;
            akk = fma(-l10, l10, d11);          // next pivot
#endif
            const double lk = a[k] * ri;
            a[k] = lk;
            if (lane < 8) colbuf[k * 8 + r] = lk;
            if (lane == k) rinvbuf[k] = ri;
            __syncwarp();
            if (k < 7) {
#pragma unroll
                for (int jj = (k + 1) & ~1; jj < 8; jj += 2) {
                    const double2 lp = *reinterpret_cast<const double2*>(colbuf + k * 8 + jj);
                    if (jj > k) { Lc[k][jj] = lp.x; a[jj] = fma(-lk, lp.x, a[jj]); }
                    Lc[k][jj + 1] = lp.y; a[jj + 1] = fma(-lk, lp.y, a[jj + 1]);
                }
            }
        }
        GP_SUB(17);
        // inverse, lane r owns column r of X = L^-1; right-looking, so the dependent chain is one fma and one mul per row
        double x[8], sacc[8];
#pragma unroll
        for (int i = 0; i < 8; i++) sacc[i] = 0.0;
#pragma unroll
        for (int k = 0; k < 8; k++) {
            const double rkk = rinvbuf[k];
            x[k] = (k == r) ? rkk : ((k > r) ? -sacc[k] * rkk : 0.0);
#pragma unroll
            for (int i = k + 1; i < 8; i++) sacc[i] = fma(Lc[k][i], x[k], sacc[i]);
        }
        GP_SUB(18);
        if (lane < 8) {
            // x[i] = D[i][r]: atom (ni = kb, kc = 2 kb + r/4), element (i, r%4); zeros above the diagonal are written too
            double* at = sm.linv + (kb * (kb + 1) + 2 * kb + (r >> 2)) * 32 + (r & 3);
#pragma unroll
            for (int k = 0; k < 8; k++) {
                if (k <= r) Cs[(c0 + r) * CS_LD + c0 + k] = a[k];
                at[k * 4] = x[k];
            }
        }
        GP_SUB(19);
    };
    // tile (ri, ci) -= X_ri X_ci^T, X = column block kb of the rows below
    auto update_tile = [&](const int kb, const int ri, const int ci) {
        const int c0 = kb * 8;
        const double* arow = Cs + (ri * 8 + g) * CS_LD + c0 + q;
        const double* brow = Cs + (ci * 8 + g) * CS_LD + c0 + q;
        double2* cp = reinterpret_cast<double2*>(Cs + (ri * 8 + g) * CS_LD + ci * 8 + 2 * q);
        const double2 cv = *cp;
        double acc[2] = {cv.x, cv.y};
        dmma(acc, neg_bits(arow[0]), brow[0]);
        dmma(acc, neg_bits(arow[4]), brow[4]);
        *cp = make_double2(acc[0], acc[1]);
    };
    if (warp_u == 0) potf2(0);
    __syncthreads();
    GP_PHASE_MARK(8);
    for (int kb = 0; kb < 8; kb++) {
        const int c0 = kb * 8;
        // X = A21 D^T for the row tiles below (tile 8 = right-hand-side rows); the result replaces A21 and, negated, fills the
        // sub-diagonal atoms of sm.linv
        {
            const int ri = kb + 1 + warp;
            if (ri <= 8) {
                double acc[2] = {0.0, 0.0};
                const double* arow = Cs + (ri * 8 + g) * CS_LD + c0 + q;
                const double* bat = sm.linv + (kb * (kb + 1) + 2 * kb) * 32 + lane;
                const double a0 = arow[0], a1 = arow[4];
                dmma(acc, a0, bat[0]);
                dmma(acc, a1, bat[32]);
                __syncwarp();   // every lane has read its A fragments before the tile is overwritten
                *reinterpret_cast<double2*>(Cs + (ri * 8 + g) * CS_LD + c0 + 2 * q) = make_double2(acc[0], acc[1]);
                if (ri < 8)
                    *reinterpret_cast<double2*>(sm.linv + (ri * (ri + 1) + 2 * kb + (q >> 1)) * 32 + g * 4 + ((2 * q) & 3)) =
                        make_double2(neg_bits(acc[0]), neg_bits(acc[1]));
            }
        }
        __syncthreads();
        GP_PHASE_MARK(9);
        // trailing update, tiles (ri, ci) with kb < ci <= min(ri, 7), ri <= 8. Warp 0 updates the next diagonal tile first and goes
        // straight on to its 8x8 factorisation (the serial part of P2) while the other warps share the remaining tiles.
        if (warp_u == 0) {
            if (kb < 7) {
                update_tile(kb, kb + 1, kb + 1);
                __syncwarp();
                potf2(kb + 1);
            }
        } else {
            int idx = warp - 1;
            for (int ri = kb + 1; ri <= 8; ri++) {
                const int first = (ri == kb + 1) ? 1 : 0;          // tile (kb+1, kb+1) belongs to warp 0
                const int width = min(ri, 7) - kb - first;
                while (idx < width) {
                    update_tile(kb, ri, kb + 1 + first + idx);
                    idx += FWARPS - 1;
                }
                idx -= width;
            }
        }
        __syncthreads();
        GP_PHASE_MARK(8);
    }
}

// ---------------------------------------------------------------------------------------------- main routine
struct Pipe { uint32_t produced; uint32_t consumed; };   // slab sequence numbers of the operand ring (produced is kept equal to
                                                         // consumed at phase boundaries; the ring itself tracks only consumed)

// Start of block row I: rows below sm.pre_split belong to the shared read-only prefix factor, the others to this task's scratch, which
// then starts at block row pre_split (pre_split == 0: one scratch holds everything).
__device__ __forceinline__ const double* row_ptr(const FactorSmem& sm, const double* scratch, int I) {
    return (I < sm.pre_split) ? sm.pre_lo + row_off(I) : scratch + (row_off(I) - row_off(sm.pre_split));
}

// Row-tile operand producer: slab t of tile `tile` of panel j (A rows of blocks I0 [, I0+1] and the B rows of block j) goes
// into pipeline slot gi, whose stage must be free. Called by one lane.
__device__ __forceinline__ void issue_row_slab(FactorSmem& sm, const double* scratch, int j, int T, int blk0, int blk_end, int tile,
                                               int t, uint32_t gi) {
    const int I0 = blk0 + 2 * tile;
    const bool two = (I0 + 1 < blk_end);
    const int st = gi % STAGES;
    mbar_expect_tx(&sm.full[st], (two ? 3 : 2) * SLAB_D * 8);
    double* dst = sm.stage + st * STAGE_D;
    const size_t so = (size_t)t * SLAB_D;
#if GPSLC_L2_HINTS
    const uint64_t pa = l2_policy_evict_first(), pb = l2_policy_evict_last();
    bulk_g2s_hint(dst, row_ptr(sm, scratch, I0) + so, SLAB_D * 8, &sm.full[st], pa);
    if (two) bulk_g2s_hint(dst + SLAB_D, row_ptr(sm, scratch, I0 + 1) + so, SLAB_D * 8, &sm.full[st], pa);
    bulk_g2s_hint(dst + 2 * SLAB_D, row_ptr(sm, scratch, j) + so, SLAB_D * 8, &sm.full[st], pb);
#else
    bulk_g2s(dst, row_ptr(sm, scratch, I0) + so, SLAB_D * 8, &sm.full[st]);
    if (two) bulk_g2s(dst + SLAB_D, row_ptr(sm, scratch, I0 + 1) + so, SLAB_D * 8, &sm.full[st]);
    bulk_g2s(dst + 2 * SLAB_D, row_ptr(sm, scratch, j) + so, SLAB_D * 8, &sm.full[st]);
#endif
}

// The k-loop of one row tile, slabs [tb, te): acc -= nothing yet, acc += A_slab x B_slab^T over the slabs. It is a separate
// NON-INLINED function on purpose: inlined into factor_run, the 64 accumulator registers plus the operand fragments
// compete with everything factor_run keeps live across the loop (generator, panel and tile bookkeeping), and ptxas
// spilled accumulators on every slab iteration; as a call, the caller's state is saved once per tile and the loop runs
// spill-free. acc travels through local memory (accio, MI*16 doubles): zero-initialised when tb == 0.
// gi0 = pipeline sequence number of slab tb; a stage is refilled (slab STAGES ahead in the panel's (tile, t) order) by the
// last warp that leaves it.
template <int MI>
__device__ __noinline__ void row_tile_kloop(double* __restrict__ accio, const double* __restrict__ scratch, const int j, const int T,
                                            const int blk0, const int blk_end, const int tile, const int F, const int tb, const int te,
                                            const uint32_t gi0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = (MI == 2) ? (warp >> 2) : 0;
    const int r8base = (MI == 2) ? (warp & 3) * 2 : warp;
    double acc[MI][8][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < 8; ni++) {
            acc[mi][ni][0] = (tb == 0) ? 0.0 : accio[(mi * 8 + ni) * 2];
            acc[mi][ni][1] = (tb == 0) ? 0.0 : accio[(mi * 8 + ni) * 2 + 1];
        }
    uint32_t gi = gi0;
#ifdef GPSLC_PHASE_TIMING
    long long _wait = 0, _prod = 0, _nref = 0;
    const long long _f0 = clock64();
#endif
    const int fbase = tile * T + STAGES;
    for (int t = tb; t < te; t++, gi++) {
        const int st = gi % STAGES;
#ifdef GPSLC_PHASE_TIMING
        const long long _w0 = clock64();
#endif
        mbar_wait(&sm.full[st], (gi / STAGES) & 1);
#ifdef GPSLC_PHASE_TIMING
        _wait += clock64() - _w0;
#endif
        const double* sA = sm.stage + st * STAGE_D + half * SLAB_D + r8base * (K4S * 32);
        const double* sB = sm.stage + st * STAGE_D + 2 * SLAB_D;
#pragma unroll
        for (int k4 = 0; k4 < K4S; k4++) {
            double a[MI];
#pragma unroll
            for (int mi = 0; mi < MI; mi++) a[mi] = sA[mi * (K4S * 32) + k4 * 32 + lane];
#pragma unroll
            for (int ni = 0; ni < 8; ni++) {
                const double b = sB[(ni * K4S + k4) * 32 + lane];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) dmma(acc[mi][ni], a[mi], b);
            }
        }
        __syncwarp();
        if (lane == 0 && stage_release_is_last(&sm.freed[st]) && fbase + t < F) {
            // this warp was the last one out of the stage: refill it with the slab STAGES further down the panel's (tile, t) order
            int pt = t + STAGES, ptile = tile;
            if (pt >= T) { pt -= T; ptile++; }
#ifdef GPSLC_PHASE_TIMING
            const long long _p0 = clock64();
#endif
            issue_row_slab(sm, scratch, j, T, blk0, blk_end, ptile, pt, gi + STAGES);
#ifdef GPSLC_PHASE_TIMING
            _prod += clock64() - _p0; _nref++;
#endif
        }
    }
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < 8; ni++) {
            accio[(mi * 8 + ni) * 2] = acc[mi][ni][0];
            accio[(mi * 8 + ni) * 2 + 1] = acc[mi][ni][1];
        }
#ifdef GPSLC_PHASE_TIMING
    if (lane == 0) {   // all warps: slots 12/13/14 are sums over the 8 warps
        atomicAdd(&g_phase_cycles[12], (unsigned long long)_wait);
        atomicAdd(&g_phase_cycles[13], (unsigned long long)(clock64() - _f0));
        atomicAdd(&g_phase_cycles[14], (unsigned long long)_prod);
        atomicAdd(&g_phase_cycles[32 + warp], (unsigned long long)(clock64() - _f0));
        atomicAdd(&g_phase_cycles[40 + warp], (unsigned long long)_wait);
        atomicAdd(&g_phase_cycles[48 + warp], (unsigned long long)_nref);
    }
#endif
}

// Diagonal-tile slot of warp `warp` (see factor_run): the 36 lower 8x8 tiles of a diagonal block are dealt 5/4 to the warps —
// warps p and p+4 share the row-tile pair (p, 7-p), which has 9 tiles: warp p takes row p and the first 4-p tiles of row 7-p,
// warp p+4 the other four. Slot sl of a warp is tile (row, col).
constexpr int DSLOTS = 5;
__device__ __forceinline__ void diag_slot(const int warp, const int sl, int& row, int& col) {
    const int pw = warp & 3;
    if (warp < 4) { row = (sl <= pw) ? pw : 7 - pw; col = (sl <= pw) ? sl : sl - pw - 1; }
    else { row = 7 - pw; col = min(4 - pw + sl, 7); }
}

// The k-loop of the FIRST row tile of panel j with the diagonal tile of the same panel folded in: the B slabs of a row tile are
// the slabs of block row j, which are both operands of C_jj -= L_jJ L_jJ^T and the matrix of the forward-solve update
// w_j -= L_jJ z_J. Folding saves the separate diagonal pass over the same slabs (its pipeline fill and drain, its barrier rounds,
// one re-read of j*32 KB) and runs the 36 diagonal tiles at the row tiles' pipe utilisation. The operand pipeline covers this
// tile only (F = T), so that the stage buffers are idle afterwards and P2 can alias them.
// accio: [MI*16] row accumulators, then [2*DSLOTS] diagonal accumulators, then [MAXRHS] partial sums of L_jJ z_J (per thread:
// row tid/4, k columns tid%4 + 4i).
template <int MI>
__device__ __noinline__ void row_tile_kloop_fold(double* __restrict__ accio, const double* __restrict__ scratch,
                                                 const double* __restrict__ zbuf, const int npad, const int nrhs, const int j, const int T,
                                                 const int blk0, const int blk_end, const uint32_t gi0) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    FactorSmem& sm = *reinterpret_cast<FactorSmem*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int half = (MI == 2) ? (warp >> 2) : 0;
    const int r8base = (MI == 2) ? (warp & 3) * 2 : warp;
    double acc[MI][8][2];
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < 8; ni++) { acc[mi][ni][0] = 0.0; acc[mi][ni][1] = 0.0; }
    double dacc[DSLOTS][2];
    int offa[DSLOTS], offb[DSLOTS];
    const int nslots = (warp < 4) ? 5 : 4;
#pragma unroll
    for (int sl = 0; sl < DSLOTS; sl++) {
        dacc[sl][0] = 0.0; dacc[sl][1] = 0.0;
        int r, c;
        diag_slot(warp, sl, r, c);
        offa[sl] = r * (K4S * 32) + lane; offb[sl] = c * (K4S * 32) + lane;
    }
    double wsum[MAXRHS] = {0.0, 0.0};
    const int wr = threadIdx.x >> 2, kq = threadIdx.x & 3;
    const int loff = (wr >> 3) * (K4S * 32) + (wr & 7) * 4;
    uint32_t gi = gi0;
    for (int t = 0; t < T; t++, gi++) {
        const int st = gi % STAGES;
        mbar_wait(&sm.full[st], (gi / STAGES) & 1);
        const double* sA = sm.stage + st * STAGE_D + half * SLAB_D + r8base * (K4S * 32);
        const double* sB = sm.stage + st * STAGE_D + 2 * SLAB_D;
#pragma unroll
        for (int k4 = 0; k4 < K4S; k4++) {
            double a[MI];
#pragma unroll
            for (int mi = 0; mi < MI; mi++) a[mi] = sA[mi * (K4S * 32) + k4 * 32 + lane];
#pragma unroll
            for (int ni = 0; ni < 8; ni++) {
                const double b = sB[(ni * K4S + k4) * 32 + lane];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) dmma(acc[mi][ni], a[mi], b);
            }
#pragma unroll
            for (int sl = 0; sl < DSLOTS; sl++)
                if (sl < nslots) dmma(dacc[sl], sB[offa[sl] + k4 * 32], sB[offb[sl] + k4 * 32]);
        }
        if (nrhs > 0) {
#pragma unroll
            for (int kk = 0; kk < KB / 4; kk++) {
                const int kc = kq + 4 * kk;
                const double l = sB[loff + (kc >> 2) * 32 + (kc & 3)];
                for (int rh = 0; rh < nrhs; rh++) wsum[rh] = fma(l, zbuf[(size_t)rh * npad + t * KB + kc], wsum[rh]);
            }
        }
        __syncwarp();
        if (lane == 0 && stage_release_is_last(&sm.freed[st]) && t + STAGES < T)
            issue_row_slab(sm, scratch, j, T, blk0, blk_end, 0, t + STAGES, gi + STAGES);
    }
#pragma unroll
    for (int mi = 0; mi < MI; mi++)
#pragma unroll
        for (int ni = 0; ni < 8; ni++) {
            accio[(mi * 8 + ni) * 2] = acc[mi][ni][0];
            accio[(mi * 8 + ni) * 2 + 1] = acc[mi][ni][1];
        }
#pragma unroll
    for (int sl = 0; sl < DSLOTS; sl++) { accio[MI * 16 + 2 * sl] = dacc[sl][0]; accio[MI * 16 + 2 * sl + 1] = dacc[sl][1]; }
    accio[MI * 16 + 2 * DSLOTS] = wsum[0];
    accio[MI * 16 + 2 * DSLOTS + 1] = wsum[1];
}

// Gen concept:
//   void quad(int r0, int r1, int c, double& v00, double& v01, double& v10, double& v11) const
//        -> K[r0][c], K[r0][c+1], K[r1][c], K[r1][c+1]   (lower triangle / rectangular rows; c even)
//   double rhs(int which, int r) const
// GPSLC_FOLD = 1 folds the diagonal tile's k-loop into the k-loop of the panel's first row tile (row_tile_kloop_fold: same B slabs,
// one pass less over them). Measured on B200 (profiles/fold_ab_r02.md): bit-identical chains, but 3 % SLOWER at c3 (1277 vs 1314
// sweeps/s), 13 % slower at c2 and 2.5 % at c4 - the folded loop issues one LDS per DMMA (the diagonal operands are not shared
// between tiles of a warp) and gives up the two-slab stages of the stand-alone diagonal loop. Kept as a build-time experiment, off.
#ifndef GPSLC_FOLD
#define GPSLC_FOLD 0
#endif
// PRE (shared prefix): the leading pre_split block rows / panels of the matrix are an already computed factor that several tasks
// share read-only (ITE: chol(Kp) does not depend on doT, so it is computed once per posterior sample and every doT value continues
// from it - src/prediction.jl:31-33 redoes it per doT). Panels j < pre_split are REPLAYED: no diagonal tile, no P2 (their outputs
// come from pre_linv), only the row tiles of this task's own rows (>= pre_split) against the shared B slabs; the caller preloads
// zbuf[0 .. pre_split*NB) with the prefix's forward solve. `scratch` then holds block rows >= pre_split only.
// save_linv (any mode): P2's outputs of every panel are also written to save_linv[j][LINV_D] - what a later PRE run replays.
template <class Gen, int TEAM = 0, bool SNAP = false, bool FOLD = (GPSLC_FOLD != 0) && !SNAP, bool PRE = false>
__device__ void factor_run(const Gen& gen, const int NRB, const int NCB, const int nrhs, double* scratch,
                           double* zbuf /* [MAXRHS][NCB*NB] solves, then [MAXRHS][NCB*NB] pre-solve w */,
                           FactorSmem& sm, Pipe& pipe, const int snapJ = 1 << 30, double* snap = nullptr, const int snap_n = 0,
                           const double* pre_lo = nullptr, const double* pre_linv = nullptr, const int pre_split = 0,
                           double* save_linv = nullptr, const bool keep_diag = true) {
    // keep_diag = false: the diagonal blocks L_jj are not written to the scratch. No k-loop ever reads them (the diagonal loop of panel
    // j' covers block row j' up to column j'-1, the row tiles use sm.linv), so callers that only want log det and the solves - every MH
    // proposal and slice evaluation of the sampler - skip 32 KB of stores per panel (5 % of a CTA's time at n = 256, 1 % at n = 1024).
    // snapshot hook (SNAP = true, ITE path only — the sampler's instantiation carries none of its code, which matters for the
    // register budget of the k-loops): for panels j >= snapJ the value K - sum_{J<snapJ} L L^T (the Schur complement of the
    // leading snapJ panels, i.e. CovITE + jitter*I) is written to snap[(c-snapJ*NB)*snap_n + (r-snapJ*NB)] (both triangles).
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, q = lane & 3;
    const int npad = NCB * NB;
    const int trank = (TEAM == 1) ? (int)cluster_rank() : (TEAM == 2) ? (int)blockIdx.x : 0;
    const int tsize = (TEAM == 1) ? (int)cluster_size() : (TEAM == 2) ? (int)gridDim.x : 1;
    if (tid == 0) {
        sm.info = 0; sm.snap = snap; sm.snapJ = snapJ; sm.snap_n = snap_n;
        sm.pre_lo = pre_lo; sm.pre_linv = pre_linv; sm.pre_split = PRE ? pre_split : 0; sm.save_linv = save_linv;
    }
    if (tid < NB) { sm.part[0][tid] = 0.0; sm.part[1][tid] = 0.0; sm.part[2][tid] = 0.0; sm.part[3][tid] = 0.0; }
    __syncthreads();

    GP_PHASE_INIT();
    for (int j = 0; j < NCB; j++) {
        const int T = j * NSLAB;  // slabs in the k-loop of this panel
        // Row blocks below the diagonal that this CTA owns. Team mode: the nblk blocks are dealt out in contiguous runs, the ranks
        // that get one block more rotate with the panel index.
        const bool replay = PRE && j < pre_split;
        int nblk = replay ? NRB - pre_split : NRB - j - 1, blk0 = replay ? pre_split : j + 1;
        if constexpr (TEAM != 0) {
            const int er = (trank + j) % tsize, per = nblk / tsize, rem = nblk - per * tsize;
            blk0 += er * per + min(er, rem);
            nblk = per + (er < rem ? 1 : 0);
        }
        int blk_end = blk0 + nblk;
        // FOLD: the diagonal tile's k-loop rides on the k-loop of this CTA's first row tile (row_tile_kloop_fold) whenever there is
        // one; the accumulators of that tile wait in accm (local memory) while P2 runs and its epilogue follows P2.
        const bool fold = FOLD && j > 0 && nblk > 0;
        const bool fold2 = fold && (blk0 + 1 < blk_end);      // the folded tile has two blocks (128 rows)
        double accm[2 * 16 + 2 * DSLOTS + MAXRHS];
        // =================================================================== diagonal tile
        if (replay) {
            // shared prefix panel: column features for the generators and the saved P2 outputs; readers of the previous panel's
            // copies are behind its end-of-panel barrier
            gen.stage_cols(j * NB, sm.colfeat);
            for (int i = tid; i < LINV_D; i += FTHREADS) sm.linv[i] = pre_linv[(size_t)j * LINV_D + i];
            __syncthreads();
        } else {
            // the 36 lower 8x8 tiles of the diagonal block are dealt 5/4 to the warps (diag_slot); slot s of a warp is tile (srow[s], scol[s])
            const int nslots = (warp < 4) ? 5 : 4;
            int srow[DSLOTS], scol[DSLOTS];
#pragma unroll
            for (int sl = 0; sl < DSLOTS; sl++) diag_slot(warp, sl, srow[sl], scol[sl]);
            // tile rows of the slots: warps 0-3 hold row pw in slots 0..pw and row 7-pw in the others, warps 4-7 row 7-pw only
            const int rowA = (warp < 4) ? (warp & 3) : 7 - (warp & 3), rowB = 7 - (warp & 3);
            const int splitA = (warp < 4) ? (warp & 3) : DSLOTS;
            double acc[DSLOTS][2];
#pragma unroll
            for (int sl = 0; sl < DSLOTS; sl++) { acc[sl][0] = 0.0; acc[sl][1] = 0.0; }
            double wsum[MAXRHS] = {0.0, 0.0};
            double wsnap[MAXRHS] = {0.0, 0.0};  // residual after the leading snapJ panels only (ITE: -MeanITE)
            const int wr = tid >> 2, kq = tid & 3;  // RHS update mapping: row wr, k pair kq
            // The diagonal k-loop needs only the block row's own slabs (8 KB each), so a stage carries DS = 2 of them (one 16 KB
            // bulk copy): half as many barrier rounds per DMMA. Pair t2 goes into pipeline slot gi (stage known to be free); the
            // first STAGES pairs are issued here, the others by the last warp that leaves the stage they reuse.
            constexpr int DS = 2;
            const int T2 = T / DS;      // T = 4 j is a multiple of 4
            auto issue_diag_slab = [&](const int t2, const uint32_t gi) {
                const int st = gi % STAGES;
                mbar_expect_tx(&sm.full[st], DS * SLAB_D * 8);
#if GPSLC_L2_HINTS
                bulk_g2s_hint(sm.stage + st * STAGE_D + SLAB_D, row_ptr(sm, scratch, j) + (size_t)t2 * DS * SLAB_D, DS * SLAB_D * 8, &sm.full[st],
                              l2_policy_evict_last());     // the row tiles of this panel read the same slabs again
#else
                bulk_g2s(sm.stage + st * STAGE_D + SLAB_D, row_ptr(sm, scratch, j) + (size_t)t2 * DS * SLAB_D, DS * SLAB_D * 8, &sm.full[st]);
#endif
            };
            if (!fold) {
                for (int t2 = 0; t2 < STAGES && t2 < T2; t2++) {
                    const uint32_t gi = pipe.consumed + t2;
                    if (lane == 0 && warp == (int)(gi & (FWARPS - 1))) issue_diag_slab(t2, gi);
                }
            } else {
                // operand pipeline over the first row tile only (its B slabs are the diagonal block row's slabs)
                for (int f = 0; f < STAGES && f < T; f++) {
                    const uint32_t gi = pipe.consumed + f;
                    if (lane == 0 && warp == (int)(gi & (FWARPS - 1))) issue_row_slab(sm, scratch, j, T, blk0, blk_end, 0, f, gi);
                }
            }
            // column features of this panel for the generators, while the first operand copies are in flight; they are first
            // read after the barrier that follows the k-loop, and the previous panel's readers are behind its end-of-panel barrier
            gen.stage_cols(j * NB, sm.colfeat);
            const bool in_tail = SNAP && (j >= snapJ);
            const bool do_snap = in_tail && (snap != nullptr) && (trank == 0);
            const int Tsnap = in_tail ? snapJ * NSLAB : T;
            auto kloop = [&](const int tb, const int te) {     // slab pairs [tb, te)
                for (int t2 = tb; t2 < te; t2++) {
                    const uint32_t gi = pipe.consumed++;
                    const int st = gi % STAGES;
                    mbar_wait(&sm.full[st], (gi / STAGES) & 1);
#pragma unroll
                    for (int kk2 = 0; kk2 < DS; kk2++) {
                        const double* sB = sm.stage + st * STAGE_D + SLAB_D + kk2 * SLAB_D;
#pragma unroll
                        for (int k4 = 0; k4 < K4S; k4++) {
                            // the slots of a warp lie in at most two tile rows (diag_slot): their A fragments are loaded once per
                            // row, not once per tile - this loop is otherwise bound by the shared-memory pipe (two LDS.64 per DMMA)
                            const double aA = sB[(rowA * K4S + k4) * 32 + lane];
                            const double aB = (warp < 4) ? sB[(rowB * K4S + k4) * 32 + lane] : aA;
#pragma unroll
                            for (int sl = 0; sl < DSLOTS; sl++) {
                                if (sl < nslots) {
                                    const double a = (sl > splitA) ? aB : aA;
                                    const double b = sB[(scol[sl] * K4S + k4) * 32 + lane];
                                    dmma(acc[sl], a, b);
                                }
                            }
                        }
                        if (nrhs > 0) {
                            // row wr of the diagonal block row, k columns kq, kq+4, ... of the slab (one double per 8x4 atom row)
                            const int t = t2 * DS + kk2;
#pragma unroll
                            for (int kk = 0; kk < KB / 4; kk++) {
                                const int kc = kq + 4 * kk;    // column inside the slab
                                const double l = sB[(wr >> 3) * (K4S * 32) + (kc >> 2) * 32 + (wr & 7) * 4 + (kc & 3)];
                                for (int rh = 0; rh < nrhs; rh++) wsum[rh] = fma(l, zbuf[(size_t)rh * npad + t * KB + kc], wsum[rh]);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0 && stage_release_is_last(&sm.freed[st]) && t2 + STAGES < T2) issue_diag_slab(t2 + STAGES, gi + STAGES);
                }
            };
            if (!fold) {
                kloop(0, Tsnap / DS);
            } else {
                if (fold2) row_tile_kloop_fold<2>(accm, scratch, zbuf, npad, nrhs, j, T, blk0, blk_end, pipe.consumed);
                else row_tile_kloop_fold<1>(accm, scratch, zbuf, npad, nrhs, j, T, blk0, blk_end, pipe.consumed);
                pipe.consumed += T;
                const int ao = fold2 ? 32 : 16;
#pragma unroll
                for (int sl = 0; sl < DSLOTS; sl++) { acc[sl][0] = accm[ao + 2 * sl]; acc[sl][1] = accm[ao + 2 * sl + 1]; }
                wsum[0] = accm[ao + 2 * DSLOTS]; wsum[1] = accm[ao + 2 * DSLOTS + 1];
            }
            if constexpr (SNAP) {
                if (in_tail) {
                    wsnap[0] = wsum[0]; wsnap[1] = wsum[1];
                    if (do_snap) {
                        // Schur complement of the leading snapJ panels for this diagonal tile (lower tiles; mirrored)
#pragma unroll
                        for (int sl = 0; sl < DSLOTS; sl++) {
                            if (sl < nslots) {
                                const int r = j * NB + srow[sl] * 8 + g;
                                const int c = j * NB + scol[sl] * 8 + 2 * q;
                                double v00, v01, v10, v11;
                                gen.quad(r, r, c, v00, v01, v10, v11);
                                const int ri = r - snapJ * NB, ci = c - snapJ * NB;
                                // diagonal 8x8 tiles are computed in full: only their lower triangle is mirrored, so that every
                                // element of the snapshot has exactly one writer and the result is exactly symmetric
                                if (ri < snap_n) {
                                    if (ci <= ri) { snap[(size_t)ci * snap_n + ri] = v00 - acc[sl][0]; snap[(size_t)ri * snap_n + ci] = v00 - acc[sl][0]; }
                                    if (ci + 1 <= ri) { snap[(size_t)(ci + 1) * snap_n + ri] = v01 - acc[sl][1]; snap[(size_t)ri * snap_n + ci + 1] = v01 - acc[sl][1]; }
                                }
                            }
                        }
                    }
                    kloop(Tsnap / DS, T2);
                }
            }
            __syncthreads();  // every warp is done with the stage buffers -> P2 may alias them
            GP_PHASE_MARK(0);
            double* Cs = sm.stage;   // [CS_ROWS][CS_LD]
            // C_jj = K_jj - acc  (lower tiles only; diagonal tiles in full)
#pragma unroll
            for (int sl = 0; sl < DSLOTS; sl++) {
                if (sl < nslots) {
                    const int rl = srow[sl] * 8 + g, cl = scol[sl] * 8 + 2 * q;
                    double v[2][1][2];
                    gen.template strip<1, true>(j * NB + rl, j * NB + rl, j * NB + cl, v, sm.colfeat, cl);
                    *reinterpret_cast<double2*>(Cs + rl * CS_LD + cl) = make_double2(v[0][0][0] - acc[sl][0], v[0][0][1] - acc[sl][1]);
                }
            }
            for (int rh = 0; rh < nrhs; rh++) {
                double s = wsum[rh];
                s += __shfl_xor_sync(0xffffffffu, s, 1);
                s += __shfl_xor_sync(0xffffffffu, s, 2);
                const double yv = gen.rhs(rh, j * NB + wr);
                if (kq == 0) { sm.wvec[rh][wr] = yv - s; Cs[(NB + rh) * CS_LD + wr] = yv - s; }
                if (in_tail) {
                    double s2 = wsnap[rh];
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 1);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, 2);
                    if (kq == 0) zbuf[(size_t)(MAXRHS + rh) * npad + j * NB + wr] = yv - s2;
                }
            }
            for (int idx = tid; idx < (CS_ROWS - NB - nrhs) * NB; idx += FTHREADS)   // zero padding below the right-hand sides
                Cs[(NB + nrhs + (idx >> 6)) * CS_LD + (idx & 63)] = 0.0;
            __syncthreads();
            GP_PHASE_MARK(1);
            p2_factor_diag(sm, Cs, j * NB);
            GP_PHASE_MARK(5);
            // store L_jj (lower, zeros above), log-diagonal, z_j = L_jj^-1 w_j (rows 64.. of the bordered factorisation)
            {
                double* dst = const_cast<double*>(row_ptr(sm, scratch, j)) + (size_t)j * BLOCK_D;
                if (trank == 0) {
                    if (keep_diag)
                    for (int idx = tid; idx < BLOCK_D; idx += FTHREADS) {
                        const int r = idx >> 6, c = idx & 63;
                        dst[elem_off(r, c)] = (c <= r) ? Cs[r * CS_LD + c] : 0.0;
                    }
                    if (save_linv)
                        for (int i = tid; i < LINV_D; i += FTHREADS) save_linv[(size_t)j * LINV_D + i] = sm.linv[i];
                }
                if (tid < NB) sm.part[0][tid] += log(Cs[tid * CS_LD + tid]);
                if (tid < NB * nrhs) {
                    const int rh = tid >> 6, r = tid & 63;
                    zbuf[(size_t)rh * npad + j * NB + r] = Cs[(NB + rh) * CS_LD + r];
                    if (!in_tail) zbuf[(size_t)(MAXRHS + rh) * npad + j * NB + r] = sm.wvec[rh][r];
                }
            }
            if (tid < NB && nrhs > 0) {
                // gram sums straight from the bordered rows of Cs
                const double z0 = Cs[NB * CS_LD + tid];
                sm.part[1][tid] = fma(z0, z0, sm.part[1][tid]);
                if (nrhs > 1) {
                    const double z1 = Cs[(NB + 1) * CS_LD + tid];
                    sm.part[2][tid] = fma(z0, z1, sm.part[2][tid]);
                    sm.part[3][tid] = fma(z1, z1, sm.part[3][tid]);
                }
            }
            fence_proxy_async();   // generic-proxy writes (smem workspace, global L_jj) before later async-proxy copies
            __syncthreads();
            GP_PHASE_MARK(6);
        }
        // =================================================================== row tiles below the diagonal
        // Blocks blk0.. are processed two at a time (128-row tiles, warp w owns rows 16w..16w+15); an odd leftover block is
        // processed as a 64-row tile (warp w owns rows 8w..8w+7) so that it costs half a tile of tensor time, not a full one.
        // slab index at which the Schur-complement snapshot is taken, or -1 (parameters live in shared memory)
        const int Tsnap = (SNAP && snap != nullptr && j >= snapJ) ? snapJ * NSLAB : -1;
        // MI = 2: 128-row tile (blocks I0, I0+1); MI = 1: 64-row tile (block I0 only). do_kloop = false: the accumulators are
        // already in accm (the folded first tile).
        auto run_tile = [&](auto mi_tag, const int tile, const bool do_kloop, const int F) {
            constexpr int MI = decltype(mi_tag)::value;
            const int I0 = blk0 + 2 * tile;
            const int half = (MI == 2) ? (warp >> 2) : 0;
            const int I = I0 + half;                                   // block this warp works on
            const int r8base = (MI == 2) ? (warp & 3) * 2 : warp;      // first 8-row group of this warp inside block I
            double acc[MI][8][2];                                      // accm (local memory) is filled by the non-inlined k-loop
            const int r0 = I * NB + r8base * 8 + g;
            auto kloop = [&](const int tb, const int te) {
                row_tile_kloop<MI>(accm, scratch, j, T, blk0, blk_end, tile, F, tb, te, pipe.consumed);
                pipe.consumed += te - tb;
            };
            auto fetch_acc = [&]() {
#pragma unroll
                for (int mi = 0; mi < MI; mi++)
#pragma unroll
                    for (int ni = 0; ni < 8; ni++) { acc[mi][ni][0] = accm[(mi * 8 + ni) * 2]; acc[mi][ni][1] = accm[(mi * 8 + ni) * 2 + 1]; }
            };
            if constexpr (SNAP) {
                const int Ts = (Tsnap >= 0) ? Tsnap : T;
                kloop(0, Ts);
                if (Tsnap >= 0) {
                    fetch_acc();
                    double* snap = sm.snap;
                    const int snapJ = sm.snapJ, snap_n = sm.snap_n;
#pragma unroll
                    for (int h4 = 0; h4 < 2; h4++) {
                        double v[2][4][2];
                        gen.template strip<4, MI == 1>(r0, r0 + 8, j * NB + h4 * 32 + 2 * q, v, sm.colfeat, h4 * 32 + 2 * q);
#pragma unroll
                        for (int nn = 0; nn < 4; nn++) {
                            const int ni = h4 * 4 + nn;
                            const int ci = j * NB + ni * 8 + 2 * q - snapJ * NB;
#pragma unroll
                            for (int mi = 0; mi < MI; mi++) {
                                const int ri = r0 + 8 * mi - snapJ * NB;
                                if (ri < snap_n) {
#pragma unroll
                                    for (int e = 0; e < 2; e++) {
                                        if (ci + e < snap_n) {
                                            const double x = v[mi][nn][e] - acc[mi][ni][e];
                                            snap[(size_t)(ci + e) * snap_n + ri] = x;
                                            snap[(size_t)ri * snap_n + ci + e] = x;
                                        }
                                    }
                                }
                            }
                        }
                    }
                    kloop(Ts, T);
                }
            } else {
                if (do_kloop) kloop(0, T);
            }
            GP_PHASE_MARK(2);
            fetch_acc();
            // C = K - acc. The row index is laundered through an empty asm so that nothing of the covariance generation
            // (feature loads, address arithmetic) can be scheduled into the k-loop, where every register is needed
            int r0g = r0;
            asm volatile("" : "+r"(r0g));
#pragma unroll
            for (int h4 = 0; h4 < 2; h4++) {
                double v[2][4][2];
                gen.template strip<4, MI == 1>(r0g, r0g + 8, j * NB + h4 * 32 + 2 * q, v, sm.colfeat, h4 * 32 + 2 * q);
#pragma unroll
                for (int nn = 0; nn < 4; nn++)
#pragma unroll
                    for (int mi = 0; mi < MI; mi++) {
                        acc[mi][h4 * 4 + nn][0] = v[mi][nn][0] - acc[mi][h4 * 4 + nn][0];
                        acc[mi][h4 * 4 + nn][1] = v[mi][nn][1] - acc[mi][h4 * 4 + nn][1];
                    }
            }
            // L_Ij = C L_jj^-T by block forward substitution over the eight 8-column tiles, right-looking:
            //   X_b = R_b D_b^T   (D_b = inverse of the b-th 8x8 diagonal block; sm.linv atoms (b, 2b), (b, 2b+1))
            //   R_c -= X_b L_cb^T for c > b   (atoms (c, 2b), (c, 2b+1) hold -L_cb)
            // so the 64x64 inverse of L_jj is never formed. A-fragments are rebuilt from accumulator-layout values with quad
            // shuffles (lane (g,q) needs V[g][4h+q], held by lane (g, 2h + q/2), element q%2).
            double* dst = const_cast<double*>(row_ptr(sm, scratch, I)) + (size_t)j * BLOCK_D;
            auto to_afrag = [&](const double (&t)[2], const int hh) {
                const int src = (lane & ~3) | (2 * hh + (q >> 1));
                const double v0 = __shfl_sync(0xffffffffu, t[0], src);
                const double v1 = __shfl_sync(0xffffffffu, t[1], src);
                return (q & 1) ? v1 : v0;
            };
#pragma unroll
            for (int b = 0; b < 8; b++) {
                const double* datom = sm.linv + (b * (b + 1) + 2 * b) * 32 + lane;
                double o[MI][2];
#pragma unroll
                for (int mi = 0; mi < MI; mi++) {
                    o[mi][0] = 0.0; o[mi][1] = 0.0;
                    const double r0f = to_afrag(acc[mi][b], 0), r1f = to_afrag(acc[mi][b], 1);
                    dmma(o[mi], r0f, datom[0]);
                    dmma(o[mi], r1f, datom[32]);
                    *reinterpret_cast<double2*>(dst + elem_off((r8base + mi) * 8 + g, b * 8 + 2 * q)) = make_double2(o[mi][0], o[mi][1]);
                }
                if (b < 7) {
                    double xf[MI][2];
#pragma unroll
                    for (int mi = 0; mi < MI; mi++) { xf[mi][0] = to_afrag(o[mi], 0); xf[mi][1] = to_afrag(o[mi], 1); }
#pragma unroll
                    for (int c = b + 1; c < 8; c++) {
                        const double* latom = sm.linv + (c * (c + 1) + 2 * b) * 32 + lane;
                        const double l0 = latom[0], l1 = latom[32];
#pragma unroll
                        for (int mi = 0; mi < MI; mi++) { dmma(acc[mi][c], xf[mi][0], l0); dmma(acc[mi][c], xf[mi][1], l1); }
                    }
                }
            }
        };
        if (fold) {
            // epilogue of the folded first tile (its k-loop ran before P2), then the remaining tiles as a panel of their own
            if (fold2) run_tile(std::integral_constant<int, 2>{}, 0, false, 0);
            else run_tile(std::integral_constant<int, 1>{}, 0, false, 0);
            blk0 = min(blk0 + 2, blk_end);
            nblk = blk_end - blk0;
        }
        {
            const int ntile = (nblk + 1) >> 1;
            const int F = ntile * T;
            // prologue of the operand pipeline: the first STAGES slabs (the k-loops issue the rest as they go)
            for (int f = 0; f < STAGES && f < F; f++) {
                const uint32_t gi = pipe.consumed + f;
                if (lane == 0 && warp == (int)(gi & (FWARPS - 1))) issue_row_slab(sm, scratch, j, T, blk0, blk_end, f / T, f % T, gi);
            }
            for (int tile = 0; tile < ntile; tile++) {
                if (blk0 + 2 * tile + 1 < blk_end) run_tile(std::integral_constant<int, 2>{}, tile, true, F);
                else run_tile(std::integral_constant<int, 1>{}, tile, true, F);
            }
        }
        pipe.produced = pipe.consumed;   // everything issued for this panel has been consumed
        GP_PHASE_MARK(3);
#ifdef GPSLC_PHASE_TIMING
        const long long _b0 = clock64();
#endif
        team_sync<TEAM>(sm.gbar);
#ifdef GPSLC_PHASE_TIMING
        if (lane == 0) atomicAdd(&g_phase_cycles[24 + warp], (unsigned long long)(clock64() - _b0));
#endif
        GP_PHASE_MARK(4);
    }
    // ---- reductions
    __syncthreads();
    if (tid < 4) {
        double acc = 0.0;
        for (int i = 0; i < NB; i++) acc += sm.part[tid][i];
        if (tid == 0) { sm.out.logdet = 2.0 * acc; sm.out.info = sm.info; }
        else sm.out.gram[tid - 1] = acc;
    }
    __syncthreads();
}

// Row i of (L x_s), s < NS, for the triangular factor held in the scratch blocks (Ioff + i/64, Joff + c/64): used for draws
// y = L z (logitT initialisation / slice proposals, ITE samples). x: [NS][ldx].
template <int NS>
__device__ __forceinline__ void tri_matvec_row(const double* scratch, int NRB, int Ioff, int Joff, int i, int ncols, const double* x,
                                               int ldx, int ns, double (&accv)[NS]) {
#pragma unroll
    for (int s = 0; s < NS; s++) accv[s] = 0.0;
    const int ib = i >> 6, ri = i & 63;
    for (int jb = 0; jb <= ib; jb++) {
        const double* blk = scratch + block_off(Ioff + ib, Joff + jb, NRB);
        const int jmax = (jb == ib) ? ri : 63;
        for (int jj = 0; jj <= jmax; jj++) {
            const int col = jb * 64 + jj;
            if (col < ncols) {
                const double l = blk[elem_off(ri, jj)];
#pragma unroll
                for (int s = 0; s < NS; s++)
                    if (s < ns) accv[s] = fma(l, x[(size_t)s * ldx + col], accv[s]);
            }
        }
    }
}

}  // namespace gpslc
