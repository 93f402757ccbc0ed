"""CPU restatement of the opacity build that precedes every MC iteration.

TEST INFRASTRUCTURE ONLY.  numpy restatement of, paths relative to /root/reference/tardis/:
    StimulatedEmissionFactor.calculate                 plasma/properties/radiative_properties.py:66-116
    calculate_sobolev_line_opacity                     opacities/tau_sobolev.py:21-75 (SOBOLEV_COEFFICIENT :9-18)
    numba_calculate_beta_sobolev                       opacities/tau_sobolev.py:77-88
    probability_emission_down / _internal_down / _internal_up
                                                       opacities/macro_atom/macroatom_line_transitions.py:94-139,213-247,324-367
    BoundBoundMacroAtomSolver.normalize_transition_probabilities
                                                       opacities/macro_atom/macroatom_solver.py:708-739
    + the row order of the probabilities table         macroatom_solver.py:425-436,524-585
Pinned by tests/golden/opacity_*.npz: tau, beta and the raw probabilities come from the UNMODIFIED reference functions
(oracle/reference_runner.py::run_reference_opacity); the stimulated-emission factor and the normalisation live in modules that
cannot be imported here (they pull the atomic-data / HDF stack), so the golden generator evaluates the same pandas / numpy
expressions next to them -- those two steps are pinned by restatement only, and the header of the golden says so."""
from __future__ import annotations

import numpy as np

# CODATA-2010 cgs (tardis/constants.py:1)
E_ESU, M_E, C, H = 4.80320425e-10, 9.10938291e-28, 2.99792458e10, 6.62606957e-27
SOBOLEV_COEFFICIENT = (np.pi * E_ESU**2) / (M_E * C)                 # tau_sobolev.py:9-18
C_EINSTEIN = 4.0 * (np.pi * E_ESU) ** 2 / (C * M_E)                  # macroatom_line_transitions.py:9-11


def stimulated_emission_factor(atomic, level_number_density, nlte=False):
    """radiative_properties.py:75-116"""
    n_lower = level_number_density.take(atomic.lower_level, axis=0)
    n_upper = level_number_density.take(atomic.upper_level, axis=0)
    g_lower = atomic.g[atomic.lower_level][np.newaxis].T
    g_upper = atomic.g[atomic.upper_level][np.newaxis].T
    meta_stable_upper = atomic.metastable[atomic.upper_level][np.newaxis].T
    stim = np.zeros(n_lower.shape, dtype=np.float64)
    mask = n_lower == 0.0
    with np.errstate(divide="ignore", invalid="ignore"):
        stim[~mask] = 1 - ((g_lower * n_upper)[~mask] / (g_upper * n_lower)[~mask])
    stim[np.isneginf(stim)] = 0.0
    stim[meta_stable_upper & (stim < 0)] = 0.0
    if nlte:
        stim[(stim < 0) & atomic.nlte_line[np.newaxis].T] = 0.0
    return stim


def tau_sobolev(atomic, level_number_density, time_explosion, stim):
    """tau_sobolev.py:55-61"""
    tau = ((atomic.wavelength_cm * atomic.f_lu)[np.newaxis].T * SOBOLEV_COEFFICIENT * time_explosion * stim
           * level_number_density.take(atomic.lower_level, axis=0))
    if np.any(np.isnan(tau)) or np.any(np.isinf(np.abs(tau))):
        raise ValueError("Some tau_sobolevs are nan, inf, -inf in tau_sobolevs. Something went wrong!")
    return tau


def beta_sobolev(tau):
    """tau_sobolev.py:77-88"""
    beta = np.empty_like(tau)
    big, small = tau > 1e3, tau < 1e-4
    mid = ~(big | small)
    beta[big] = tau[big] ** -1
    beta[small] = 1 - 0.5 * tau[small]
    beta[mid] = (1 - np.exp(-tau[mid])) / tau[mid]
    return beta


def macro_atom_probabilities(atomic, beta, stim, j_blues):
    """raw probabilities per macro-atom row (macroatom_line_transitions.py:129-139,242-247,356-367), then normalised per
    source block with NaN -> 0 (macroatom_solver.py:731-739)"""
    nu = atomic.nu[:, None]
    e_lower = atomic.energy[atomic.lower_level][:, None]
    e_upper = atomic.energy[atomic.upper_level][:, None]
    f_ul, f_lu = atomic.f_ul[:, None], atomic.f_lu[:, None]
    p_emission = beta * (2 * nu**2 * f_ul / C**2 * (e_upper - e_lower)) * C_EINSTEIN
    p_down = beta * (2 * nu**2 * f_ul / C**2 * e_lower) * C_EINSTEIN
    p_up = beta * (f_lu / (H * nu) * stim * j_blues * e_lower) * C_EINSTEIN
    rows = atomic.transition_line_idx
    raw = np.where((atomic.transition_type == -1)[:, None], p_emission[rows],
                   np.where((atomic.transition_type == 0)[:, None], p_down[rows], p_up[rows]))
    starts = atomic.macro_block_edge_index[:-1]
    sums = np.add.reduceat(raw, starts, axis=0)
    with np.errstate(divide="ignore", invalid="ignore"):
        norm = raw / sums[atomic.source_block]
    norm[np.isnan(norm)] = 0.0
    return raw, norm


def build(atomic, plasma, nlte=False):
    stim = stimulated_emission_factor(atomic, plasma.level_number_density, nlte)
    tau = tau_sobolev(atomic, plasma.level_number_density, plasma.time_explosion, stim)
    beta = beta_sobolev(tau)
    raw, norm = macro_atom_probabilities(atomic, beta, stim, plasma.j_blues)
    return dict(stimulated_emission_factor=stim, tau_sobolev=tau, beta_sobolev=beta, raw_probabilities=raw, transition_probabilities=norm)
