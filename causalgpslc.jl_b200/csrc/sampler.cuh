// Many-chain GP-SLC posterior sampler state (device tables + per-chain buffers).
// Restates src/model.jl, src/model_likelihood.jl, src/model_prior.jl, src/proposal.jl and src/inference.jl of the
// reference as data tables (which factor a hyperparameter feeds, which features a factor uses) so that one set of
// kernels serves all eight `Posterior` methods.
#pragma once
#include <vector>
#include "context.cuh"
#include "gens.cuh"
#include "rng.cuh"

namespace gpslc {

// prior families in the order of gpslc_prior.shape[] / scale[]
enum PriorFam { P_UNOISE = 0, P_XNOISE, P_TNOISE, P_YNOISE, P_XSCALE, P_TSCALE, P_YSCALE, P_UXLS, P_UTLS, P_XTLS, P_UYLS,
                P_XYLS, P_TYLS, P_NFAM };

enum SrcKind { SRC_U = 0, SRC_X = 1, SRC_T = 2 };
enum TargetKind { TGT_XCOL = 0, TGT_T = 1, TGT_LOGIT = 2, TGT_Y = 3 };

struct FactorDef {
    int exists, D;
    int src_kind[DMAX], src_idx[DMAX], ls_param[DMAX];
    int scale_param, noise_param;
    int target_kind, target_idx;
};

struct SiteDef {
    int param;     // index into the packed hyperparameter vector
    int factor;    // GP factor it feeds, or -1 (uNoise: only the U prior terms change)
    double pshape, pscale;  // InvGamma prior of this hyperparameter
};

struct ModelDev {
    int n, nU, nX, nF, binary, n_params, stride, n_obj, has_xmodel, u_layout_reference, ess_rule;
    double eps, cov, dU, drift;
    const double *X, *T, *Y;
    size_t xstride, tstride;        // per-chain data offsets (0: one dataset shared by all chains)
    const int *obj_start, *obj_of;
    const FactorDef* fdef;
    const SiteDef* sites;
    int n_sites;
    const int *lane_sites, *lane_off, *lane_factor;
    int n_lanes;
    unsigned long long seed;
    int chain0, n_chains;
    int nMH, nES;
    int dense_u;                    // SigmaU is an unstructured dense matrix: SL holds its Cholesky factor (factor.cuh scratch layout)
    const double* SL;
    int ls_unsquared;               // experiment knob (GPSLC_LS_UNSQUARED=1): kernel exp(-d^2 / ls) instead of exp(-d^2 / ls^2), see DESIGN.md §5
};

struct ChainDev {
    double *theta, *U, *Ueff, *UeffP, *Uprop, *nu, *lp, *lpP, *q, *qP, *ess, *logitT, *Xmodel;
    double* zs;                     // dense SigmaU: [C][npad] forward-solve / normal-draw scratch
    double* nuL;                    // binary T: [C][nES][n] slice directions L_stale z_j drawn once per outer iteration (App. B6)
    int *info, *infoP;
    int *active_a, *active_b;       // double-buffered compacted lists of chains still slicing
    unsigned int* n_active;         // [2]
    unsigned long long *accepts, *ess_evals, *ess_evals_logit;
};

struct Sampler {
    Ctx* ctx = nullptr;
    ModelDev m{};
    ChainDev c{};
    std::vector<std::pair<void*, size_t>> owned;   // device allocations (pointer, bytes): returned to the context's block cache
    int outer_done = 0;             // outer iterations completed so far
    int sweeps_done = 0;            // MH sweeps completed inside the current outer iteration (bench stepping)
    double* samples = nullptr;      // device [nOuter_cap][C][stride]
    int samples_cap = 0;
    size_t samples_bytes = 0;
    std::vector<int> lane_task_order;  // lanes sorted by decreasing work
    int* d_lane_order = nullptr;
    int n_exist = 0;                // number of existing GP factors
    int* d_exist = nullptr;         // their ids
    double* xibuf = nullptr;        // per-slot normal draws for L z products (binary T)
    // host copies of tables
    std::vector<FactorDef> h_fdef;
    std::vector<SiteDef> h_sites;
    std::vector<int> h_lane_sites, h_lane_off, h_lane_factor;
};

}  // namespace gpslc
