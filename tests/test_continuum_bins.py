"""Global-bin formulation of the bound-free opacity and estimators (tardis_b200/csrc/continuum_bins.cuh) against the
reference's per-continuum formulas (opacities/opacities.py:89-246, radfield_estimator_calcs.py:57-124), on the CPU.

The header is the one the CUDA kernels compile; tests/continuum_bins_shim.cpp drives it the way tb200_set_model and the
kernels do.  Bar: chi_bf_tot to 1e-13 relative, the four floating-point estimators to 1e-11, the statistics exactly."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from tardis_b200 import synthetic as syn

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
H, KB = 6.62606957e-27, 1.3806488e-16


@pytest.fixture(scope="module")
def shim():
    out = os.path.join(HERE, "_shim", "libcontinuum_bins_shim.so")
    src = os.path.join(HERE, "continuum_bins_shim.cpp")
    hdr = os.path.join(ROOT, "tardis_b200", "csrc", "continuum_bins.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.shim_continuum_bins.restype = C.c_int
    lib.shim_continuum_bins.argtypes = [C.c_int, C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.c_longlong] + [C.c_void_p] * 8
    return lib


def literal(cont, t_e, nu, energy, distance, shell):
    """chi_bf_tot and the five estimators exactly as the reference computes them, trace by trace."""
    refs = cont.photo_ion_block_references
    nc, S = len(cont.bf_threshold_list_nu), cont.chi_bf.shape[1]
    chi = np.zeros(len(nu))
    est = np.zeros((5, nc, S))
    for i in range(len(nu)):
        s = shell[i]
        bz = np.exp(-(H * nu[i]) / (KB * t_e[s]))
        for k in range(nc):
            if not (cont.photo_ion_nu_threshold_mins[k] <= nu[i] <= cont.photo_ion_nu_threshold_maxs[k]):
                continue
            a, b = refs[k], refs[k + 1]
            pn = cont.phot_nus[a:b]
            idx = np.searchsorted(pn, nu[i])
            hi, lo = a + idx, a + idx - 1 if idx > 0 else b - 1
            interval = cont.phot_nus[hi] - cont.phot_nus[lo]
            hw, lw = nu[i] - cont.phot_nus[lo], cont.phot_nus[hi] - nu[i]
            chi[i] += (cont.chi_bf[hi, s] * hw + cont.chi_bf[lo, s] * lw) / interval
            xs = (cont.x_sect[hi] * hw + cont.x_sect[lo] * lw) / interval
            inc = energy[i] * distance[i] * xs / nu[i]
            bfh = energy[i] * distance[i] * xs * (1 - cont.bf_threshold_list_nu[k] / nu[i])
            est[0, k, s] += inc
            est[1, k, s] += inc * bz
            est[2, k, s] += bfh
            est[3, k, s] += bfh * bz
            est[4, k, s] += 1.0
    return chi, est


def run_bins(shim, cont, t_e, nu, energy, distance, shell):
    nc, S, n_phot = len(cont.bf_threshold_list_nu), cont.chi_bf.shape[1], len(cont.phot_nus)
    chi = np.zeros(len(nu)); tie = np.zeros(len(nu), dtype=np.int32); bins = np.zeros(len(nu), dtype=np.int32)
    est = np.zeros((5, nc, S))
    arrs = [np.ascontiguousarray(a) for a in (cont.phot_nus, cont.photo_ion_block_references.astype(np.int64),
                                              cont.photo_ion_nu_threshold_mins, cont.photo_ion_nu_threshold_maxs, cont.x_sect,
                                              cont.chi_bf, cont.bf_threshold_list_nu, t_e)]
    tr = [np.ascontiguousarray(a) for a in (nu, energy, distance, shell.astype(np.int64))]
    rc = shim.shim_continuum_bins(n_phot, nc, S, *(a.ctypes.data for a in arrs), len(nu), *(a.ctypes.data for a in tr),
                                  chi.ctypes.data, tie.ctypes.data, bins.ctypes.data, est.ctypes.data)
    assert rc == 0
    return chi, tie, bins, est


@pytest.mark.parametrize("seed,n_continua,points", [(1, 30, (12, 30)), (2, 7, (2, 5)), (3, 60, (20, 50))])
def test_bins_match_the_per_continuum_formulas(shim, seed, n_continua, points):
    model = syn.make_model(6, 200, "macroatom", seed=seed)
    syn.add_continuum(model, seed=seed + 100, n_continua=n_continua, points=points)
    cont = model.continuum
    rng = np.random.default_rng(seed)
    n = 4000
    nu = np.exp(rng.uniform(np.log(1e14), np.log(6e16), n))  # below, inside and above every block
    energy = rng.uniform(0.5, 2.0, n) / n
    distance = 10 ** rng.uniform(12, 15, n)
    shell = rng.integers(0, model.n_shells, n)
    chi_ref, est_ref = literal(cont, model.t_electrons, nu, energy, distance, shell)
    chi, tie, bins, est = run_bins(shim, cont, model.t_electrons, nu, energy, distance, shell)
    assert tie.sum() == 0
    assert np.array_equal(bins, np.searchsorted(np.sort(cont.phot_nus), nu, side="left"))
    assert np.array_equal(chi == 0.0, chi_ref == 0.0)  # same active / inactive pattern
    np.testing.assert_allclose(chi, chi_ref, rtol=1e-13, atol=0)
    assert np.array_equal(est[4], est_ref[4])
    for q in range(4):
        np.testing.assert_allclose(est[q], est_ref[q], rtol=1e-11, atol=0)
    assert np.array_equal(est == 0.0, est_ref == 0.0)  # untouched cells stay exactly zero, like the reference's


def test_breakpoint_ties_are_flagged_for_the_literal_path(shim):
    model = syn.make_model(4, 100, "macroatom", seed=5)
    syn.add_continuum(model, seed=9, n_continua=8, points=(3, 6))
    cont = model.continuum
    nu = np.concatenate([cont.phot_nus[::3], np.nextafter(cont.phot_nus[::3], np.inf)])
    n = len(nu)
    chi, tie, bins, est = run_bins(shim, cont, model.t_electrons, nu, np.ones(n), np.ones(n), np.zeros(n, dtype=np.int64))
    assert tie[: n // 2].all() and not tie[n // 2:].any()
    # ties contribute nothing to the moments (the kernels route them through the per-continuum functions)
    chi_ref, est_ref = literal(cont, model.t_electrons, nu[n // 2:], np.ones(n // 2), np.ones(n // 2), np.zeros(n // 2, dtype=np.int64))
    assert np.array_equal(est[4], est_ref[4])
    np.testing.assert_allclose(chi[n // 2:], chi_ref, rtol=1e-13, atol=0)


def test_unexpected_table_structure_disables_the_bins(shim):
    model = syn.make_model(4, 100, "macroatom", seed=5)
    syn.add_continuum(model, seed=9, n_continua=5, points=(3, 6))
    cont = model.continuum
    cont.photo_ion_nu_threshold_mins = cont.photo_ion_nu_threshold_mins * 1.01  # not the first point of the block any more
    nc, S, n_phot = 5, 4, len(cont.phot_nus)
    z = np.zeros(1); zi = np.zeros(1, dtype=np.int32); zl = np.zeros(1, dtype=np.int64)
    arrs = [np.ascontiguousarray(a) for a in (cont.phot_nus, cont.photo_ion_block_references.astype(np.int64),
                                              cont.photo_ion_nu_threshold_mins, cont.photo_ion_nu_threshold_maxs, cont.x_sect,
                                              cont.chi_bf, cont.bf_threshold_list_nu, model.t_electrons)]
    est = np.zeros((5, nc, S))
    rc = shim.shim_continuum_bins(n_phot, nc, S, *(a.ctypes.data for a in arrs), 0, z.ctypes.data, z.ctypes.data, z.ctypes.data,
                                  zl.ctypes.data, z.ctypes.data, zi.ctypes.data, zi.ctypes.data, est.ctypes.data)
    assert rc == 1
