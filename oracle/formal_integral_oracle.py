"""CPU parity oracle of the formal integral (TEST INFRASTRUCTURE ONLY: tests/, bench.py's cpu_baseline leg).

Restates what `FormalIntegralSolver.solve` does between the source function and the spectrum
(/root/reference/tardis/spectrum/formal_integral/formal_integral_solver.py):
  :208-232  the radii of the interpolated shells (np.linspace of `interpolate_shells` points; 0 -> max(2 S, 80); < 0 -> none)
  :305-430  interpolate_integrator_quantities: scipy interp1d over the shell mid-points -- "nearest" for the electron densities and
            tau_sobolev, linear with extrapolation for att_S_ul / Jred_lu / Jblue_lu, negative values clipped to 0.  scipy is the
            reference's own dependency for this step and is called here with the reference's arguments
  :255-257  Fortran-order flattening (index = shell * n_lines + line)
  :268-277  the integrator: oracle/formal_integral_oracle.c (numba_formal_integral restated in C), then
            8 pi^2 np.trapezoid(I_nu_p, dx = r_max / n_p) per frequency (formal_integral_numba.py:559-563)
Pinned against goldens from the unmodified reference (tests/golden/formal_integral_*.npz)."""
from __future__ import annotations

import ctypes as C
import time

import numpy as np

from . import cpu_oracle

SIGMA_THOMSON = 6.652458734e-25  # transport/montecarlo/configuration/constants.py:3


def interpolated_radii(r_inner, r_outer, interpolate_shells):
    r_inner, r_outer = np.asarray(r_inner, dtype=np.float64), np.asarray(r_outer, dtype=np.float64)
    if interpolate_shells == 0:
        interpolate_shells = max(2 * len(r_inner), 80)  # formal_integral_solver.py:208-214
    if interpolate_shells > 0:
        radius = np.linspace(r_inner[0], r_outer[-1], interpolate_shells)  # :222-228
        return radius[:-1], radius[1:]
    return r_inner, r_outer  # :229-232


def interpolate_integrator_quantities(r_inner, r_outer, r_inner_i, r_outer_i, att_S_ul, Jred_lu, Jblue_lu, tau_sobolev, electron_densities):
    """formal_integral_solver.py:305-430, same calls in the same order."""
    from scipy.interpolate import interp1d

    r_middle = (np.asarray(r_inner) + np.asarray(r_outer)) / 2.0
    r_middle_i = (np.asarray(r_inner_i) + np.asarray(r_outer_i)) / 2.0
    ne_i = interp1d(r_middle, np.asarray(electron_densities, dtype=np.float64), fill_value="extrapolate", kind="nearest")(r_middle_i)
    tau_i = interp1d(r_middle, np.asarray(tau_sobolev), fill_value="extrapolate", kind="nearest")(r_middle_i)
    att_i = interp1d(r_middle, att_S_ul, fill_value="extrapolate")(r_middle_i)
    jred_i = interp1d(r_middle, Jred_lu, fill_value="extrapolate")(r_middle_i)
    jblue_i = interp1d(r_middle, Jblue_lu, fill_value="extrapolate")(r_middle_i)
    return att_i.clip(0.0), jred_i.clip(0.0), jblue_i.clip(0.0), tau_i, ne_i


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


def integrate(r_inner, r_outer, time_explosion, line_list_nu, inner_temperature, frequencies, att_S_ul, Jred_lu, Jblue_lu, tau_sobolev,
              electron_densities, n_impact_parameters, sigma_thomson=SIGMA_THOMSON):
    """numba_formal_integral (formal_integral_numba.py:377-567) on [L, S] tables of the integrator's shells.
    -> (luminosity_densities [n], intensities_nu_p [n, n_impact_parameters])."""
    lib = cpu_oracle.lib()
    f = lib.tb_oracle_formal_integral
    f.restype = C.c_int
    r_inner = np.ascontiguousarray(r_inner, dtype=np.float64)
    r_outer = np.ascontiguousarray(r_outer, dtype=np.float64)
    nu_lines = np.ascontiguousarray(line_list_nu, dtype=np.float64)
    freq = np.ascontiguousarray(frequencies, dtype=np.float64)
    S, L, P = len(r_inner), len(nu_lines), int(n_impact_parameters)
    pad = lambda t: np.concatenate([np.asarray(t, dtype=np.float64).flatten(order="F"), [0.0]])  # noqa: E731 (see the C file's header)
    att, jred, jblue = pad(att_S_ul), pad(Jred_lu), pad(Jblue_lu)
    exp_tau = np.exp(-np.asarray(tau_sobolev, dtype=np.float64).T.ravel())  # :240
    ne = np.ascontiguousarray(electron_densities, dtype=np.float64)
    inup = np.zeros((len(freq), P))
    t0 = time.perf_counter()
    rc = f(C.c_int64(S), _p(r_inner), _p(r_outer), C.c_double(time_explosion), C.c_int64(L), _p(nu_lines), C.c_double(inner_temperature),
           C.c_int64(len(freq)), _p(freq), _p(att), _p(jred), _p(jblue), _p(exp_tau), _p(ne), C.c_double(sigma_thomson), C.c_int64(P), _p(inup))
    integrate.last_c_seconds = time.perf_counter() - t0  # the integrator alone (bench.py's cpu_baseline), without the table preparation above
    if rc != 0:
        raise MemoryError("tb_oracle_formal_integral")
    radius_max = r_outer[-1]
    lum = np.array([8 * np.pi * np.pi * np.trapezoid(inup[k], dx=radius_max / P) for k in range(len(freq))])  # :559-563
    return lum, inup


def solve(r_inner, r_outer, time_explosion, line_list_nu, inner_temperature, frequencies, att_S_ul, Jred_lu, Jblue_lu, tau_sobolev,
          electron_densities, points, interpolate_shells, sigma_thomson=SIGMA_THOMSON):
    """What FormalIntegralSolver.solve computes after the source function: -> dict(luminosity_densities, intensities_nu_p, + the
    interpolated tables)."""
    r_in_i, r_out_i = interpolated_radii(r_inner, r_outer, interpolate_shells)
    att_i, jred_i, jblue_i, tau_i, ne_i = interpolate_integrator_quantities(r_inner, r_outer, r_in_i, r_out_i, att_S_ul, Jred_lu, Jblue_lu,
                                                                           tau_sobolev, electron_densities)
    lum, inup = integrate(r_in_i, r_out_i, time_explosion, line_list_nu, inner_temperature, frequencies, att_i, jred_i, jblue_i, tau_i, ne_i,
                          points, sigma_thomson)
    return dict(luminosity_densities=lum, intensities_nu_p=inup, att_S_ul_interpolated=att_i, Jred_lu_interpolated=jred_i,
                Jblue_lu_interpolated=jblue_i, tau_sobolevs_interpolated=tau_i, electron_densities_interpolated=ne_i,
                r_inner_interpolated=np.asarray(r_in_i), r_outer_interpolated=np.asarray(r_out_i))
