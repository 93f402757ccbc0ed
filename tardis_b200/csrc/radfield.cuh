// radfield.cuh -- estimator -> radiation field solve on the device (first half).
//
// What it replaces (paths relative to /root/reference/tardis/):
//   MCRadiationFieldPropertiesSolver.solve                              transport/montecarlo/estimators/mc_rad_field_solver.py:37-90
//     .estimate_dilute_planck_radiation_field (T_rad, W per shell)      :92-113
//     .estimate_jblues (normalisation, zero fill, optical window)       :115-144
//   DilutePlanckianRadiationField.calculate_mean_intensity              plasma/radiation_field/planck_rad_field.py:55-71
//   intensity_black_body                                                util/base.py:279-302
// called by Simulation.advance_state right after every MC iteration (simulation/base.py:281-288) on the estimators the
// transport kernels have just left in HBM: J, nu_bar [S] and J_blue [L, S] (80 MB at L = 5e5, S = 20) never go to the host.
//
// Plain IEEE arithmetic in the reference's operation order, shared by the kernel and the host (TBR_HD) so that it can
// be unit-tested on a CPU build of this header (tests/radfield_shim.cpp).
#pragma once

#if defined(__CUDACC__)
#define TBR_HD __host__ __device__ __forceinline__
#else
#define TBR_HD inline
#include <cmath>
#endif

namespace tbr {

struct Constants {  // computed by the caller exactly as the reference's module does (mc_rad_field_solver.py:20-30, util/base.py:21-23,299-300)
    double t_radiative_estimator_constant;  // (pi**4 / (15 * 24 * zeta(5, 1))) * (h / k_B)
    double sigma_sb, c, h, k_b;
};

// estimate_dilute_planck_radiation_field, one shell
TBR_HD void dilute_planck(const Constants &K, double j, double nu_bar, double time_of_simulation, double volume, double *t_rad, double *w) {
    const double t = K.t_radiative_estimator_constant * nu_bar / j;
    *t_rad = t;
    *w = j / (4 * K.sigma_sb * pow(t, 4.0) * time_of_simulation * volume);
}

// W * intensity_black_body(nu, T): coefficient * nu**3 / (exp(h_cgs * nu * beta_rad) - 1)
TBR_HD double dilute_planck_intensity(const Constants &K, double nu, double t_rad, double w) {
    const double beta_rad = 1 / (K.k_b * t_rad);
    const double coefficient = 2 * K.h / (K.c * K.c);
    return w * (coefficient * (nu * nu * nu) / (exp(K.h * nu * beta_rad) - 1));
}

// estimate_jblues, one (line, shell) cell.  norm = c t_exp / (4 pi t_sim V) of the shell.
TBR_HD double j_blue_cell(const Constants &K, double estimator, double norm, double nu, double t_rad, double w, double w_epsilon,
                          bool detailed_optical_window) {
    double j = estimator * norm;
    const bool zero = (j == 0.0);
    if (detailed_optical_window) {
        // line_list_wavs = (nu Hz).to(AA): c / nu * 1e8; keep the estimate only inside (2500, 10000) Angstrom
        const double wav = K.c / nu * 1e8;
        if (!(wav > 2500.0 && wav < 10000.0)) j = dilute_planck_intensity(K, nu, t_rad, w);
    }
    if (zero) j = w_epsilon * dilute_planck_intensity(K, nu, t_rad, w);
    return j;
}

}  // namespace tbr
