// Argument block of the ITE / SATE kernels (estimation.cu), shared with the C-ABI layer.
#pragma once
namespace gpslc {
struct EstArgs {
    int n, nX, nU, n_params, stride;      // record layout
    const double* X; const double* T; const double* Y;
    const double* samples;                // [n_outer][n_chains][stride]
    int n_chains;
    const int* ret_idx; int R;            // retained outer indices (0-based)
    const double* doT; int n_doT;
    double jitter; int spp;
    unsigned long long seed; int chain0;
    double* mean_out;    // [n_doT][C][R][n]
    double* cov_out;     // [n_doT][C][R][n][n]
    double* ite_out;     // [n_doT][C][R*spp][n]
    int* info;           // [n_doT][C][R]
    // SATE
    double* msate; double* vsate;   // [n_doT][C][R]
    double* sate_out;               // [n_doT][C][R*spp]
    int var_as_std;
    int ls_unsquared;               // experiment knob, see sampler.cuh
    int dot0;                       // global index of doT[0] (sharded sweeps): only the draws' RNG streams depend on it
};

struct Ctx;
int launch_summarize(Ctx* ctx, const double* samples, int batch, int m, int n, double ci, double* out);

}  // namespace gpslc
