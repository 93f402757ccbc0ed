"""Opacity build on the device (tardis_b200/csrc/opacity_build.cuh, tb200_build_opacity).

CPU: (1) the numpy oracle (oracle/opacity_oracle.py) against golden vectors computed by the reference's own functions
(calculate_sobolev_line_opacity, numba_calculate_beta_sobolev, probability_*); (2) the PRODUCT's functions -- the header the
CUDA kernels compile, built for the host by tests/opacity_build_shim.cpp -- against the same.
GPU: the kernels through the C-ABI against the goldens, and a transport run on device-built tables == the same run on the
host-built tables of the oracle.
Bar: stimulated-emission factor and tau 1e-14; beta and the probabilities 1e-11 (the reference's own (1 - exp(-tau)) / tau
amplifies an ulp of exp by 1 / tau <= 1e4 next to its series switch at tau = 1e-4: numba's and numpy's exp already differ
by 1e-12 there); zero patterns identical."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from golden_util import GOLDEN_DIR, make_golden

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CASES = list(make_golden.OPACITY_CASES)
TOL = {"stimulated_emission_factor": 1e-14, "tau_sobolev": 1e-14, "beta_sobolev": 1e-11, "raw_probabilities": 1e-11,
       "transition_probabilities": 1e-11}


def load(name):
    model, atomic, plasma = make_golden.opacity_inputs(name)
    return model, atomic, plasma, dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def check(got, want, keys=None):
    for k in keys or TOL:
        a, b = got[k], want[k]
        assert np.array_equal(a == 0, b == 0), f"{k}: zero pattern differs"
        np.testing.assert_allclose(a, b, rtol=TOL[k], atol=0, err_msg=k)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    from oracle import opacity_oracle

    model, atomic, plasma, g = load(name)
    check(opacity_oracle.build(atomic, plasma, nlte=True), g)


@pytest.fixture(scope="module")
def shim():
    out = os.path.join(HERE, "_shim", "libopacity_build_shim.so")
    src = os.path.join(HERE, "opacity_build_shim.cpp")
    hdr = os.path.join(ROOT, "tardis_b200", "csrc", "opacity_build.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.shim_opacity_build.restype = None
    lib.shim_opacity_build.argtypes = [C.c_int] * 5 + [C.c_void_p] * 16 + [C.c_double] + [C.c_void_p] * 6

    def run(atomic, plasma):
        from oracle import opacity_oracle as oo

        L, S, T = len(atomic.nu), plasma.level_number_density.shape[1], len(atomic.transition_type)
        out = {k: np.empty((L, S)) for k in ("stimulated_emission_factor", "tau_sobolev", "beta_sobolev")}
        out["raw_probabilities"], out["transition_probabilities"] = np.empty((T, S)), np.empty((T, S))
        k = np.array([oo.SOBOLEV_COEFFICIENT, oo.C_EINSTEIN, oo.C, oo.H])
        arrs = [np.ascontiguousarray(a) for a in (
            atomic.lower_level, atomic.upper_level, atomic.g, atomic.metastable.astype(np.uint8), atomic.nlte_line.astype(np.uint8),
            atomic.wavelength_cm * atomic.f_lu, atomic.f_lu, atomic.f_ul, atomic.energy[atomic.lower_level], atomic.energy[atomic.upper_level],
            atomic.nu, atomic.transition_type, atomic.transition_line_idx, atomic.macro_block_edge_index, plasma.level_number_density,
            plasma.j_blues)]
        lib.shim_opacity_build(L, atomic.n_levels, S, T, len(atomic.macro_block_edge_index) - 1, *(a.ctypes.data for a in arrs),
                               plasma.time_explosion, k.ctypes.data, out["stimulated_emission_factor"].ctypes.data, out["tau_sobolev"].ctypes.data,
                               out["beta_sobolev"].ctypes.data, out["raw_probabilities"].ctypes.data, out["transition_probabilities"].ctypes.data)
        return out

    return run


@pytest.mark.parametrize("name", CASES)
def test_product_functions_match_reference_golden(shim, name):
    model, atomic, plasma, g = load(name)
    check(shim(atomic, plasma), g)


# ---- at the bench's size: 5e5 lines, 3000 levels, 1.5e6 macro-atom rows (four shells) -------------------------------------------
def check_compressed(got, g):
    """the golden holds every table as per-shell bucket sums + 8000 cells + its number of zero cells (make_golden.compress_table)"""
    for k, tol in TOL.items():
        c = make_golden.compress_table(got[k])
        assert int(c["n_zero"]) == int(g[f"{k}__n_zero"]), f"{k}: number of zero cells differs"
        assert np.array_equal(c["sample_idx"], g[f"{k}__sample_idx"])
        assert np.array_equal(c["sample_val"] == 0, g[f"{k}__sample_val"] == 0), f"{k}: zero pattern of the sampled cells differs"
        np.testing.assert_allclose(c["sample_val"], g[f"{k}__sample_val"], rtol=tol, atol=0, err_msg=f"{k} sampled cells")
        # a bucket sums ~5000 cells (of both signs where tau is negative): the cell's bar against the largest cell, summed
        np.testing.assert_allclose(c["bucket_sums"], g[f"{k}__bucket_sums"], rtol=tol, atol=5200 * tol * float(g[f"{k}__max_abs"]), err_msg=f"{k} bucket sums")


def load_bench_shape():
    model, atomic, plasma = make_golden.opacity_inputs("opacity_bench_shape")
    return atomic, plasma, dict(np.load(os.path.join(GOLDEN_DIR, "opacity_bench_shape.npz")))


def test_oracle_matches_reference_at_the_bench_size():
    from oracle import opacity_oracle

    atomic, plasma, g = load_bench_shape()
    check_compressed(opacity_oracle.build(atomic, plasma, nlte=True), g)


def test_product_functions_match_reference_at_the_bench_size(shim):
    atomic, plasma, g = load_bench_shape()
    check_compressed(shim(atomic, plasma), g)


def engine_with_atomic(model, atomic, mode):
    from tardis_b200.engine import Engine

    eng = Engine(0)
    eng.set_option("keep_opacity_tables", 1)
    eng.set_model(r_inner=model.r_inner, r_outer=model.r_outer, time_explosion=model.time_explosion,
                  electron_density=model.electron_density, line_list_nu=model.line_list_nu, tau_sobolev=None,
                  line_interaction_type=mode, transition_probabilities=None, line2macro_level_upper=atomic.line2macro_level_upper,
                  macro_block_edge_index=atomic.macro_block_edge_index, transition_type=atomic.transition_type,
                  destination_level_id=atomic.destination_level_id, transition_line_id=atomic.transition_line_idx,
                  spectrum_frequency_grid=model.spectrum_frequency_grid)
    eng.set_atomic_data(lines_lower_level_index=atomic.lower_level, lines_upper_level_index=atomic.upper_level, g=atomic.g,
                        metastability=atomic.metastable, wavelength_cm=atomic.wavelength_cm, f_lu=atomic.f_lu, f_ul=atomic.f_ul,
                        energy_lower=atomic.energy[atomic.lower_level], energy_upper=atomic.energy[atomic.upper_level],
                        nlte_line=atomic.nlte_line)
    return eng


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_builds_the_reference_tables(name):
    model, atomic, plasma, g = load(name)
    mode = make_golden.OPACITY_CASES[name][4]
    eng = engine_with_atomic(model, atomic, mode)
    from tardis_b200.engine import EngineError

    with pytest.raises(EngineError):  # the tables are pending until build_opacity has run
        eng.upload_packets(np.ones(1) * model.r_inner[0], np.ones(1) * 1e15, np.ones(1) * 0.5, np.ones(1), np.ones(1, dtype=np.int64))
        eng.transport(True)
    eng.build_opacity(plasma.level_number_density, plasma.time_explosion, plasma.j_blues)
    got = eng.download_opacity(transition_probabilities=True)
    check(got, g, keys=("stimulated_emission_factor", "tau_sobolev", "beta_sobolev", "transition_probabilities"))
    eng.close()


@pytest.mark.gpu
def test_transport_on_device_built_tables_equals_host_built_tables(oracle):
    """The whole point: the MC iteration runs on tables that never existed on the host.  Physical (positive) populations
    here: the transport needs tau >= 0 and every activated block normalised."""
    from oracle import opacity_oracle
    from tardis_b200 import synthetic as syn

    model = syn.make_model(8, 4000, "macroatom", mu_tau=-4.0, seed=61)
    atomic = syn.make_atomic_data(model.line_list_nu, 300, "macroatom", seed=62, nlte_fraction=0.0)
    plasma = syn.make_plasma_state(atomic, 8, model.time_explosion, seed=63, zero_fraction=0.0, inversion_fraction=0.0, noise=0.0)
    plasma.level_number_density *= 1e-9  # optical depths of order one
    ref_tables = opacity_oracle.build(atomic, plasma)
    assert (ref_tables["tau_sobolev"] >= 0).all()
    packets = syn.make_packets(20000, model.r_inner[0], base_seed=12)
    host = syn.Model(r_inner=model.r_inner, r_outer=model.r_outer, v_inner=model.v_inner, v_outer=model.v_outer,
                     time_explosion=model.time_explosion, electron_density=model.electron_density, t_electrons=model.t_electrons,
                     line_list_nu=model.line_list_nu, tau_sobolev=ref_tables["tau_sobolev"],
                     macro=syn.MacroAtomTables(ref_tables["transition_probabilities"], atomic.line2macro_level_upper,
                                               atomic.macro_block_edge_index, atomic.transition_type, atomic.destination_level_id,
                                               atomic.transition_line_idx),
                     spectrum_frequency_grid=model.spectrum_frequency_grid, line_interaction_type="macroatom")
    ref = oracle.run_oracle(host, packets, nthreads=4)
    eng = engine_with_atomic(model, atomic, "macroatom")
    eng.build_opacity(plasma.level_number_density, plasma.time_explosion, plasma.j_blues)
    res = eng.run_packets(packets)
    # the device-built tables differ from numpy's by an ulp of exp here and there: same trajectories except at exact ties
    assert sum(res["counters"][k] != v for k, v in ref["counters"].items()) == 0
    np.testing.assert_allclose(res["output_nus"], ref["output_nus"], rtol=1e-9)
    for k in ("j", "nu_bar"):
        np.testing.assert_allclose(res[k], ref[k], rtol=1e-9)
    eng.close()
