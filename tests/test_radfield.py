"""Estimator -> radiation field solve (tardis_b200/csrc/radfield.cuh, tb200_solve_radiation_field).

CPU: (1) the numpy oracle (oracle/radfield_oracle.py) against golden vectors of the unmodified
`MCRadiationFieldPropertiesSolver.solve`; (2) the PRODUCT's functions -- the very header the CUDA kernels compile, built for
the host by tests/radfield_shim.cpp -- against the same, and against the oracle for the optical-window branch.
GPU: the kernels through the C-ABI, with host-supplied estimators (goldens) and on the estimators a transport run left in HBM.
Bar: 1e-13 relative (one or two ulp of exp / pow between libm, numpy and CUDA); zero / non-zero pattern identical."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from golden_util import GOLDEN_DIR, make_golden

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
RTOL = 1e-13
CASES = list(make_golden.RADFIELD_CASES)


def load(name):
    return make_golden.radfield_inputs(name), dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def check(got, want, rtol=RTOL):
    for k, (a, b) in {"t_radiative": (got[0], want["t_radiative"]), "dilution_factor": (got[1], want["dilution_factor"]),
                      "j_blues": (got[2], want["j_blues"])}.items():
        np.testing.assert_allclose(a, b, rtol=rtol, atol=0, err_msg=k)


@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    from oracle import radfield_oracle

    inp, g = load(name)
    check(radfield_oracle.solve(**inp), g, rtol=1e-15)


@pytest.fixture(scope="module")
def shim():
    out = os.path.join(HERE, "_shim", "libradfield_shim.so")
    src = os.path.join(HERE, "radfield_shim.cpp")
    hdr = os.path.join(ROOT, "tardis_b200", "csrc", "radfield.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.shim_radfield.restype = None
    lib.shim_radfield.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_double] * 3 + [C.c_int] + [C.c_void_p] * 4

    def run(inp, window=False):
        from oracle import radfield_oracle as ro

        S, L = len(inp["j"]), len(inp["line_list_nu"])
        k = np.array([ro.T_RADIATIVE_ESTIMATOR_CONSTANT, ro.SIGMA_SB, ro.C, ro.H, ro.K_B])
        t, w, jb = np.empty(S), np.empty(S), np.empty((L, S))
        a = [np.ascontiguousarray(inp[x]) for x in ("j", "nu_bar", "j_blue", "volume", "line_list_nu")]
        lib.shim_radfield(S, L, *(x.ctypes.data for x in a), inp["time_explosion"], inp["time_of_simulation"], inp["w_epsilon"],
                          int(window), k.ctypes.data, t.ctypes.data, w.ctypes.data, jb.ctypes.data)
        return t, w, jb

    return run


@pytest.mark.parametrize("name", CASES)
def test_product_functions_match_reference_golden(shim, name):
    inp, g = load(name)
    check(shim(inp), g)


# ---- at the bench's [5e5 lines, 20 shells] -----------------------------------------------------------------------------------
def check_bench_shape(got, g, rtol):
    np.testing.assert_allclose(got[0], g["t_radiative"], rtol=rtol, atol=0)
    np.testing.assert_allclose(got[1], g["dilution_factor"], rtol=rtol, atol=0)
    c = make_golden.compress_table(got[2])
    assert int(c["n_zero"]) == int(g["j_blues__n_zero"]) and np.array_equal(c["sample_idx"], g["j_blues__sample_idx"])
    np.testing.assert_allclose(c["sample_val"], g["j_blues__sample_val"], rtol=rtol, atol=0)
    np.testing.assert_allclose(c["bucket_sums"], g["j_blues__bucket_sums"], rtol=max(rtol, 1e-13), atol=0)  # all cells >= 0: no cancellation


def test_oracle_and_product_match_reference_at_the_bench_size(shim):
    from oracle import radfield_oracle

    inp = make_golden.radfield_inputs("radfield_bench_shape")
    g = dict(np.load(os.path.join(GOLDEN_DIR, "radfield_bench_shape.npz")))
    check_bench_shape(radfield_oracle.solve(**inp), g, rtol=1e-15)
    check_bench_shape(shim(inp), g, rtol=RTOL)


@pytest.mark.parametrize("name", CASES)
def test_product_optical_window_matches_oracle(shim, name):
    from oracle import radfield_oracle

    inp, _ = load(name)
    want = radfield_oracle.solve(**inp, detailed_optical_window=True)
    got = shim(inp, window=True)
    for a, b in zip(got, want):
        np.testing.assert_allclose(a, b, rtol=RTOL, atol=0)
    nu = inp["line_list_nu"]
    wav = radfield_oracle.C / nu * 1e8
    assert ((wav > 2500) & (wav < 10000)).any() and (~((wav > 2500) & (wav < 10000))).any()


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_solves_host_estimators_like_the_reference(name):
    from tardis_b200 import synthetic as syn
    from tardis_b200.engine import Engine

    inp, g = load(name)
    S, L = len(inp["j"]), len(inp["line_list_nu"])
    model = syn.make_model(S, L, "scatter", seed=3)
    model.line_list_nu = inp["line_list_nu"]
    eng = Engine(0)
    eng.set_model_from(model)
    got = eng.solve_radiation_field(time_explosion=inp["time_explosion"], time_of_simulation=inp["time_of_simulation"], volume=inp["volume"],
                                    w_epsilon=inp["w_epsilon"], estimators=(inp["j"], inp["nu_bar"], inp["j_blue"]))
    check(got, g)
    assert np.array_equal(got[2] == 0, g["j_blues"] == 0)
    eng.close()


@pytest.mark.gpu
@pytest.mark.parametrize("window", [False, True])
def test_engine_solves_the_resident_estimators_of_a_run(window):
    """After a transport run the estimators are in HBM; the solve on them == the oracle's solve on their downloaded copies."""
    from oracle import radfield_oracle
    from tardis_b200 import synthetic as syn
    from tardis_b200.engine import Engine

    model = syn.make_model(12, 6000, "macroatom", mu_tau=-4.0, seed=8)
    packets = syn.make_packets(30000, model.r_inner[0], base_seed=4)
    eng = Engine(0)
    eng.set_model_from(model)
    res = eng.run_packets(packets)
    volume = 4.0 / 3.0 * np.pi * (model.r_outer**3 - model.r_inner**3)
    t_sim = packets.time_of_simulation
    got = eng.solve_radiation_field(time_explosion=model.time_explosion, time_of_simulation=t_sim, volume=volume,
                                    detailed_optical_window=window)
    want = radfield_oracle.solve(res["j"], res["nu_bar"], res["j_blue"], model.time_explosion, t_sim, volume, model.line_list_nu,
                                 detailed_optical_window=window)
    for a, b in zip(got, want):
        np.testing.assert_allclose(a, b, rtol=RTOL, atol=0)
    assert (res["j_blue"] == 0).any()  # the zero fill was exercised
    eng.close()


@pytest.mark.gpu
def test_host_mirror_solver_uses_resident_or_given_estimators():
    """MCRadiationFieldPropertiesSolverB200.solve (the reference's signature): on the arrays the last run returned it works
    on their resident copies, on any other arrays it uploads them -- same numbers either way, and the oracle's."""
    from oracle import radfield_oracle
    from tardis_b200 import montecarlo as mc
    from tardis_b200 import synthetic as syn

    model = syn.make_model(8, 3000, "downbranch", mu_tau=-4.0, seed=21)
    packets = syn.make_packets(20000, model.r_inner[0], base_seed=6)
    pc = mc.PacketCollection(packets.initial_radii, packets.initial_nus, packets.initial_mus, packets.initial_energies,
                             packets.packet_seeds, packets.radiation_field_luminosity)
    geometry = mc.HomologousGeometry(model.r_inner, model.r_outer, model.v_inner, model.v_outer, model.time_explosion)
    cfg = mc.MonteCarloConfiguration()
    cfg.LINE_INTERACTION_TYPE = 1
    opacity = mc.OpacityState.from_model(model)
    vhist, vtracker, bulk, line = mc.montecarlo_transport_with_vpackets(
        pc, geometry, model.time_explosion, opacity, cfg, model.spectrum_frequency_grid, mc.generate_tracker_last_interaction_list(len(packets)), 0)
    solver = mc.MCRadiationFieldPropertiesSolverB200(1e-10)
    volume = geometry.volume
    a = solver.solve(bulk, line, model.time_explosion, packets.time_of_simulation, volume, model.line_list_nu)
    bulk2 = mc.EstimatorsBulk(bulk.mean_intensity_total.copy(), bulk.mean_frequency.copy())
    line2 = mc.EstimatorsLine(line.mean_intensity_blueward.copy(), line.energy_deposition_line_rate.copy())
    b = solver.solve(bulk2, line2, model.time_explosion, packets.time_of_simulation, volume, model.line_list_nu)
    want = radfield_oracle.solve(bulk.mean_intensity_total, bulk.mean_frequency, line.mean_intensity_blueward, model.time_explosion,
                                 packets.time_of_simulation, volume, model.line_list_nu)
    for got in (a, b):
        st = got.dilute_blackbody_radiationfield_state
        np.testing.assert_allclose(np.asarray(getattr(st.temperature, "value", st.temperature)), want[0], rtol=RTOL)
        np.testing.assert_allclose(st.dilution_factor, want[1], rtol=RTOL)
        np.testing.assert_allclose(got.j_blues, want[2], rtol=RTOL)
