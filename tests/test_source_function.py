"""Formal-integral source function on the device (tardis_b200/csrc/source_function.cuh,
tb200_solve_source_function, tardis_b200/source_function.py).

CPU: (1) the numpy / scipy oracle (oracle/source_function_oracle.py) against golden vectors computed by the reference's own
SourceFunctionSolver.solve; (2) the PRODUCT's functions -- the header the CUDA kernels compile, built for the host by
tests/source_function_shim.cpp, with the same loops as the kernels -- against the same.
GPU: the kernels through the C-ABI against the goldens (estimators given on the host), on the estimators a transport run left
in HBM against the oracle on the downloaded ones, and the host mirror with the reference's pandas frames.
Bar: identical zero patterns and |difference| <= 1e-11 |reference| + 1e-14 max|table|.  The reference solves (I - Q)^T C = e
with a sparse LU, the product iterates C <- e + Q^T C to a relative change of 1e-15 (same unique solution).  Both carry the
cancellation of the reference's `1 - exp(-tau)` (one ulp of exp becomes 1e-16 / tau of a line's term), and with the mildly
inverted populations of the goldens (tau < 0) a level's e_dot_u is a sum of terms of both signs: entries that cancel to 1e-5 of
their terms differ by 1e-11 between numpy's and libm's / CUDA's exp -- hence the absolute part of the bar."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from golden_util import GOLDEN_DIR, make_golden

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CASES = list(make_golden.SOURCE_FUNCTION_CASES)
KEYS = ("att_S_ul", "Jred_lu", "Jblue_lu", "e_dot_u")
RTOL = 1e-11
C_CGS = 2.99792458e10


def load(name):
    return make_golden.source_function_inputs(name), dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def check(got, want, rtol=RTOL):
    for k in KEYS:
        a, b = np.asarray(got[k]), want[k]
        assert a.shape == b.shape, k
        assert np.array_equal(a == 0, b == 0), f"{k}: zero pattern differs"
        np.testing.assert_allclose(a, b, rtol=rtol, atol=1e-14 * np.max(np.abs(b)), err_msg=k)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    from oracle import source_function_oracle as sfo

    i, g = load(name)
    got = sfo.solve(i["atomic"], i["tau_sobolev"], i["transition_probabilities"], i["j_blue_estimator"], i["e_dot_lu_estimator"],
                    i["time_explosion"], i["time_of_simulation"], i["volume"], i["mode"])
    assert np.array_equal(got["e_dot_u_levels"], g["e_dot_u_levels"])
    check(got, g, rtol=1e-13)


@pytest.fixture(scope="module")
def shim():
    out = os.path.join(HERE, "_shim", "libsource_function_shim.so")
    src = os.path.join(HERE, "source_function_shim.cpp")
    hdr = os.path.join(ROOT, "tardis_b200", "csrc", "source_function.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.shim_source_function.restype = C.c_int
    lib.shim_source_function.argtypes = ([C.c_int] * 5 + [C.c_void_p] * 10 + [C.c_double] * 4 + [C.c_int] + [C.c_void_p] * 4)
    return lib


def run_shim(lib, i, tol=1e-15, max_it=100000):
    a = i["atomic"]
    L, S = i["tau_sobolev"].shape
    T = len(a.transition_type)
    out = {k: np.empty((L, S)) for k in ("att_S_ul", "Jred_lu", "Jblue_lu")}
    out["e_dot_u"] = np.empty((a.n_levels, S))
    arrs = [np.ascontiguousarray(x) for x in (a.lower_level, a.upper_level, a.transition_type, a.transition_line_idx, a.wavelength_cm,
                                              i["tau_sobolev"], i["transition_probabilities"], i["j_blue_estimator"], i["e_dot_lu_estimator"],
                                              i["volume"])]
    it = lib.shim_source_function(L, S, T, a.n_levels, int(i["mode"] == "macroatom"), *(x.ctypes.data for x in arrs), i["time_explosion"],
                                  i["time_of_simulation"], C_CGS, tol, max_it, out["att_S_ul"].ctypes.data, out["Jred_lu"].ctypes.data,
                                  out["Jblue_lu"].ctypes.data, out["e_dot_u"].ctypes.data)
    return out, it


@pytest.mark.parametrize("name", CASES)
def test_product_functions_match_reference_golden(shim, name):
    i, g = load(name)
    out, it = run_shim(shim, i)
    assert it >= 0 and (it == 0) == (i["mode"] == "downbranch")
    out["e_dot_u"] = out["e_dot_u"][g["e_dot_u_levels"]]
    check(out, g)


# ---- at the bench's size: 5e5 lines, 3000 levels, 1e6 internal macro-atom rows (four shells; every shell is its own system) --------
def load_bench_shape():
    return make_golden.source_function_inputs("source_function_bench_shape"), dict(np.load(os.path.join(GOLDEN_DIR, "source_function_bench_shape.npz")))


def check_compressed(got, g, rtol):
    """the golden holds e_dot_u in full and every [L, S] table as per-shell bucket sums + 8000 cells (make_golden.compress_table)"""
    a, b = np.asarray(got["e_dot_u"]), g["e_dot_u"]
    assert a.shape == b.shape and np.array_equal(a == 0, b == 0)
    np.testing.assert_allclose(a, b, rtol=rtol, atol=1e-14 * np.max(np.abs(b)), err_msg="e_dot_u")
    for k in ("att_S_ul", "Jred_lu", "Jblue_lu"):
        c = make_golden.compress_table(got[k])
        top = float(g[f"{k}__max_abs"])
        assert int(c["n_zero"]) == int(g[f"{k}__n_zero"]), f"{k}: number of zero cells differs"
        assert np.array_equal(c["sample_idx"], g[f"{k}__sample_idx"])
        np.testing.assert_allclose(c["sample_val"], g[f"{k}__sample_val"], rtol=rtol, atol=1e-14 * top, err_msg=f"{k} sampled cells")
        # a bucket sums ~5000 cells: the bar of one cell, summed
        np.testing.assert_allclose(c["bucket_sums"], g[f"{k}__bucket_sums"], rtol=rtol, atol=5200 * 1e-14 * top, err_msg=f"{k} bucket sums")


def test_oracle_matches_reference_at_the_bench_size():
    from oracle import source_function_oracle as sfo

    i, g = load_bench_shape()
    got = sfo.solve(i["atomic"], i["tau_sobolev"], i["transition_probabilities"], i["j_blue_estimator"], i["e_dot_lu_estimator"],
                    i["time_explosion"], i["time_of_simulation"], i["volume"], i["mode"])
    assert np.array_equal(got["e_dot_u_levels"], g["e_dot_u_levels"])
    check_compressed(got, g, rtol=1e-13)


def test_product_functions_match_reference_at_the_bench_size(shim):
    """the kernels' loops (warp-ordered group sums, Jacobi sweeps over the CSR-by-destination list) on the CPU at full size"""
    i, g = load_bench_shape()
    out, it = run_shim(shim, i)
    assert it > 0
    out["e_dot_u"] = out["e_dot_u"][g["e_dot_u_levels"]]
    check_compressed(out, g, rtol=RTOL)


def test_fixed_point_reports_failure_instead_of_a_wrong_answer(shim):
    """Too few sweeps allowed: the solver must say so (the engine turns this into an error code), not return an unconverged table."""
    i, g = load("source_function_macroatom")
    out, it = run_shim(shim, i, max_it=2)
    assert it == -1


def engine_for(i):
    from tardis_b200.engine import Engine

    a, m = i["atomic"], i["model"]
    eng = Engine(0)
    eng.set_option("keep_opacity_tables", 1)
    eng.set_model(r_inner=m.r_inner, r_outer=m.r_outer, time_explosion=m.time_explosion, electron_density=m.electron_density,
                  line_list_nu=m.line_list_nu, tau_sobolev=i["tau_sobolev"], line_interaction_type=i["mode"],
                  transition_probabilities=i["transition_probabilities"], line2macro_level_upper=a.line2macro_level_upper,
                  macro_block_edge_index=a.macro_block_edge_index, transition_type=a.transition_type,
                  destination_level_id=a.destination_level_id, transition_line_id=a.transition_line_idx,
                  spectrum_frequency_grid=m.spectrum_frequency_grid)
    return eng


def solve_on(eng, i, estimators):
    a = i["atomic"]
    return eng.solve_source_function(time_explosion=i["time_explosion"], time_of_simulation=i["time_of_simulation"], volume=i["volume"],
                                     wavelength_cm=a.wavelength_cm, lines_lower_level_idx=a.lower_level, lines_upper_level_idx=a.upper_level,
                                     n_levels=a.n_levels, estimators=estimators)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_matches_reference_golden(name):
    i, g = load(name)
    eng = engine_for(i)
    got = solve_on(eng, i, (i["j_blue_estimator"], i["e_dot_lu_estimator"]))
    assert (got["iterations"] == 0) == (i["mode"] == "downbranch")
    got["e_dot_u"] = got["e_dot_u"][g["e_dot_u_levels"]]
    check(got, g)
    eng.close()


@pytest.mark.gpu
def test_engine_without_kept_tables_fails_loudly():
    from tardis_b200.engine import Engine, EngineError

    i, g = load("source_function_downbranch")
    a, m = i["atomic"], i["model"]
    eng = Engine(0)
    eng.set_model(r_inner=m.r_inner, r_outer=m.r_outer, time_explosion=m.time_explosion, electron_density=m.electron_density,
                  line_list_nu=m.line_list_nu, tau_sobolev=i["tau_sobolev"], line_interaction_type=i["mode"],
                  transition_probabilities=i["transition_probabilities"], line2macro_level_upper=a.line2macro_level_upper,
                  macro_block_edge_index=a.macro_block_edge_index, transition_type=a.transition_type,
                  destination_level_id=a.destination_level_id, transition_line_id=a.transition_line_idx,
                  spectrum_frequency_grid=m.spectrum_frequency_grid)
    with pytest.raises((EngineError, ValueError)):
        solve_on(eng, i, (i["j_blue_estimator"], i["e_dot_lu_estimator"]))
    eng.close()


@pytest.mark.gpu
def test_source_function_of_resident_estimators_and_host_mirror(oracle):
    """The real sequence: a Monte Carlo iteration leaves J_blue / Edotlu in HBM, the solver reads them there; the host mirror takes
    the reference's own arguments (pandas frames of atom_data / MacroAtomState, state namespaces)."""
    import types

    import pandas as pd

    from oracle import opacity_oracle
    from oracle import source_function_oracle as sfo
    from tardis_b200 import synthetic as syn
    from tardis_b200.source_function import SourceFunctionSolverB200

    S, L, n_levels = 8, 4000, 300
    model = syn.make_model(S, L, "macroatom", mu_tau=-4.0, seed=61)
    atomic = syn.make_atomic_data(model.line_list_nu, n_levels, "macroatom", seed=62, nlte_fraction=0.0)
    plasma = syn.make_plasma_state(atomic, S, model.time_explosion, seed=63, zero_fraction=0.0, inversion_fraction=0.0, noise=0.0)
    plasma.level_number_density *= 1e-9
    tables = opacity_oracle.build(atomic, plasma)
    volume = 4.0 / 3.0 * np.pi * (model.r_outer**3 - model.r_inner**3)
    i = dict(model=model, atomic=atomic, tau_sobolev=tables["tau_sobolev"], transition_probabilities=tables["transition_probabilities"],
             volume=volume, time_explosion=float(model.time_explosion), time_of_simulation=2.0e5, mode="macroatom")
    eng = engine_for(i)
    packets = syn.make_packets(50000, model.r_inner[0], base_seed=12)
    res = eng.run_packets(packets)
    assert np.count_nonzero(res["edotlu"]) > 1000
    want = sfo.solve(atomic, i["tau_sobolev"], i["transition_probabilities"], res["j_blue"], res["edotlu"], i["time_explosion"],
                     i["time_of_simulation"], volume, "macroatom")
    got = solve_on(eng, i, None)  # estimators: where the transport left them
    assert got["iterations"] > 0
    got_rows = dict(got)
    got_rows["e_dot_u"] = got["e_dot_u"][want["e_dot_u_levels"]]
    check(got_rows, want)
    # host mirror with the reference's argument objects
    ns = types.SimpleNamespace
    lines = pd.DataFrame({"line_id": np.arange(L), "wavelength_cm": atomic.wavelength_cm},
                         index=pd.MultiIndex.from_arrays([np.full(L, 14), np.full(L, 1), atomic.lower_level, atomic.upper_level],
                                                         names=["atomic_number", "ion_number", "level_number_lower", "level_number_upper"]))
    refs = pd.Series(np.arange(n_levels), index=pd.MultiIndex.from_arrays([np.full(n_levels, 14), np.full(n_levels, 1), np.arange(n_levels)],
                                                                          names=["atomic_number", "ion_number", "level_number"]))
    state = SourceFunctionSolverB200("macroatom", eng).solve(
        ns(geometry=ns(v_inner_boundary_idx=0, v_outer_boundary_idx=S), no_of_shells=S, volume=volume, time_explosion=i["time_explosion"]),
        ns(tau_sobolev=i["tau_sobolev"], transition_probabilities=i["transition_probabilities"]),
        ns(estimators_line=ns(mean_intensity_blueward=res["j_blue"], energy_deposition_line_rate=res["edotlu"]),
           packet_collection=ns(time_of_simulation=i["time_of_simulation"])),
        ns(lines=lines), ns(references_index=refs))
    assert np.array_equal(state.att_S_ul, got["att_S_ul"]) and np.array_equal(state.Jred_lu, got["Jred_lu"])
    assert list(state.e_dot_u.index.names) == ["atomic_number", "ion_number", "source_level_number"]
    assert np.array_equal(np.asarray([ix[2] for ix in state.e_dot_u.index]), want["e_dot_u_levels"])
    np.testing.assert_allclose(state.e_dot_u.to_numpy(), want["e_dot_u"], rtol=RTOL, atol=1e-14 * np.max(np.abs(want["e_dot_u"])))
    with pytest.raises(ValueError):
        SourceFunctionSolverB200("scatter", eng)
    eng.close()
