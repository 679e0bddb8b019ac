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
    // shared-prefix mode (several doT values per posterior sample): chol(Kp) of base task b = chain*R + r was computed once by
    // ite_base_kernel; this launch covers the base tasks [base0, base0 + base_n) and all doT values
    int base0, base_n;
    const double* base_L;           // [base_n][slot_lo] factor scratch of Kp (rows < NCB1)
    const double* base_linv;        // [base_n][NCB1][LINV_D]
    const double* base_z;           // [base_n][npad] L11^-1 Y
    const int* base_info;           // [base_n]
    size_t slot_lo;
};

struct Ctx;
int launch_summarize(Ctx* ctx, const double* samples, int batch, int m, int n, double ci, double* out);
int launch_subset_mean(Ctx* ctx, const double* samples, const unsigned char* mask, size_t rows, int m, int n, int groups, int n_d,
                       int count, double* out);

}  // namespace gpslc
