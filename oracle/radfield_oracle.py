"""CPU restatement of the estimator -> radiation field solve (first half).

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py's checker legs).  numpy restatement of, paths relative to
/root/reference/tardis/:
    MCRadiationFieldPropertiesSolver.solve / .estimate_dilute_planck_radiation_field / .estimate_jblues
        transport/montecarlo/estimators/mc_rad_field_solver.py:20-30 (constants), :37-144
    DilutePlanckianRadiationField.calculate_mean_intensity   plasma/radiation_field/planck_rad_field.py:55-71
    intensity_black_body                                      util/base.py:279-302 (numexpr expression evaluated with numpy)
Pinned by tests/golden/radfield_*.npz, produced by the UNMODIFIED reference class (oracle/reference_runner.py::
run_reference_radfield, detailed_optical_window=False; the window branch converts units through astropy, which is absent
here, so it is pinned by this restatement only)."""
from __future__ import annotations

import numpy as np
from scipy.special import zeta

H, K_B, C, SIGMA_SB = 6.62606957e-27, 1.3806488e-16, 2.99792458e10, 5.670373e-5  # CODATA-2010 cgs (tardis/constants.py:1)
T_RADIATIVE_ESTIMATOR_CONSTANT = (np.pi**4 / (15 * 24 * zeta(5, 1))) * (H / K_B)  # mc_rad_field_solver.py:27-29


def intensity_black_body(nu, temperature):
    """util/base.py:299-302"""
    beta_rad = 1 / (K_B * temperature)
    coefficient = 2 * H / C**2
    return coefficient * nu**3 / (np.exp(H * nu * beta_rad) - 1)


def solve(j, nu_bar, j_blue, time_explosion, time_of_simulation, volume, line_list_nu, w_epsilon=1e-10,
          detailed_optical_window=False):
    """-> (t_radiative[S], dilution_factor[S], j_blues[L,S])"""
    j, nu_bar, j_blue, volume = (np.asarray(a, dtype=np.float64) for a in (j, nu_bar, j_blue, volume))
    t_rad = T_RADIATIVE_ESTIMATOR_CONSTANT * nu_bar / j                                   # :98-102
    w = j / (4 * SIGMA_SB * t_rad**4 * time_of_simulation * volume)                      # :103-109
    norm = C * time_explosion / (4 * np.pi * time_of_simulation * volume)                # :125-129
    j_blues = j_blue * norm                                                              # :130
    planck = w * intensity_black_body(np.asarray(line_list_nu)[np.newaxis].T, t_rad)     # :131-133, planck_rad_field.py:69-71
    zero = j_blues == 0.0                                                                # :134
    if detailed_optical_window:                                                          # :135-141
        wav = C / np.asarray(line_list_nu) * 1e8
        optical = np.logical_and(wav > 2500.0, wav < 10000.0)
        j_blues[~optical] = planck[~optical]
    j_blues[zero] = w_epsilon * planck[zero]                                             # :142
    return t_rad, w, j_blues
