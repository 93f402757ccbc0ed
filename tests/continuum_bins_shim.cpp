// CPU build of tardis_b200/csrc/continuum_bins.cuh for the unit tests (tests/test_continuum_bins.py): the SAME functions the
// CUDA kernels and tb200_set_model run, driven the way they drive them.  Test harness only.
#include "../tardis_b200/csrc/continuum_bins.cuh"

// chi_out[i] = chi_bf_tot(nu[i], shell[i]) from the bin tables; tie_out[i] = 1 where the literal path would be taken.
// est_out = [5][n_continua][n_shells] estimators from the moments of all traces.  chi_bf is [n_phot][n_shells] (reference layout).
extern "C" int shim_continuum_bins(int n_phot, int n_continua, int n_shells, const double *phot_nus, const long long *refs,
                                   const double *pi_min, const double *pi_max, const double *x_sect, const double *chi_bf,
                                   const double *thr, const double *t_e, long long n_traces, const double *nu, const double *energy,
                                   const double *distance, const long long *shell, double *chi_out, int *tie_out, int *bin_out,
                                   double *est_out) {
    using namespace tbc;
    const HostBins h = build_bins(phot_nus, n_phot, refs, n_continua, pi_min, pi_max);
    if (!h.usable) return 1;
    std::vector<int> refs32(n_continua + 1);
    for (int k = 0; k <= n_continua; k++) refs32[k] = (int)refs[k];
    // per-shell rows of chi_bf, then the {C, D} tables (continuum_lin_kernel)
    std::vector<double> chi_t((size_t)n_shells * n_phot), CD((size_t)n_shells * (n_phot + 1) * 2);
    for (int s = 0; s < n_shells; s++)
        for (int i = 0; i < n_phot; i++) chi_t[(size_t)s * n_phot + i] = chi_bf[(size_t)i * n_shells + s];
    for (int s = 0; s < n_shells; s++)
        for (int g = 0; g <= n_phot; g++)
            bin_chi_linear(g, h.B.data(), n_phot, phot_nus, h.pos.data(), refs32.data(), n_continua, &chi_t[(size_t)s * n_phot],
                           &CD[((size_t)s * (n_phot + 1) + g) * 2], &CD[((size_t)s * (n_phot + 1) + g) * 2 + 1]);
    BinView v{h.B.data(), h.guide.data(), h.gkey_min, h.n_gkeys, n_phot};
    std::vector<double> mom((size_t)n_shells * (n_phot + 1) * N_MOMENTS, 0.0);
    const double H = 6.62606957e-27, KB = 1.3806488e-16;
    for (long long i = 0; i < n_traces; i++) {
        bool tie;
        const int g = find_bin(v, nu[i], &tie);
        const int s = (int)shell[i];
        tie_out[i] = tie ? 1 : 0; bin_out[i] = g;
        const double left = g >= 1 ? h.B[g - 1] : 0.0;
        const double *cd = &CD[((size_t)s * (n_phot + 1) + g) * 2];
        chi_out[i] = cd[0] + cd[1] * (nu[i] - left);
        if (tie) continue;
        const double bz = exp(-(H * nu[i]) / (KB * t_e[s]));
        const TraceMoments t = trace_moments(nu[i], energy[i], distance[i], bz, left);
        double *m = &mom[((size_t)s * (n_phot + 1) + g) * N_MOMENTS];
        for (int q = 0; q < 7; q++) m[q] += t.m[q];
    }
    const size_t ncs = (size_t)n_continua * n_shells;
    for (int k = 0; k < n_continua; k++)
        for (int s = 0; s < n_shells; s++) {
            double out[5] = {0, 0, 0, 0, 0};
            continuum_estimators_from_moments(k, &mom[(size_t)s * (n_phot + 1) * N_MOMENTS], h.B.data(), phot_nus, x_sect, h.pos.data(),
                                              refs32.data(), thr[k], out);
            for (int q = 0; q < 5; q++) est_out[q * ncs + (size_t)k * n_shells + s] = out[q];
        }
    return 0;
}
