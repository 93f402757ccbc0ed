"""Host mirror of SpectrumSolver / TARDISSpectrum (tardis_b200/spectrum.py; SURVEY.md §8f rank 2, the hand-off of the fused
histograms to the spectrum solver).  CPU: the oracle produces the packets a transport run would; the fused histograms are what the
engine's epilogue sums (numpy.histogram's bin rule -- tests/test_gpu_parity.py checks the kernel against exactly that)."""
import types
import warnings

import numpy as np
import pytest

from tardis_b200 import montecarlo as mc
from tardis_b200 import spectrum as sp
from tardis_b200 import synthetic as syn
from tardis_b200.formal_integral import IntegrationError

ns = types.SimpleNamespace


def state_from_oracle(oracle, fused=True, vpackets=0, full_relativity=False):
    model = syn.make_model(8, 3000, "macroatom", mu_tau=-4.0, seed=11)
    packets = syn.make_packets(4000, model.r_inner[0], base_seed=3)
    res = oracle.run_oracle(model, packets, number_of_vpackets=vpackets, nthreads=2)
    pc = mc.PacketCollection(packets.initial_radii, packets.initial_nus, packets.initial_mus, packets.initial_energies, packets.packet_seeds, 2.5e-5)
    pc.output_nus[:], pc.output_energies[:] = res["output_nus"], res["output_energies"]
    st = mc.MonteCarloTransportState(pc, None, None, model.time_explosion)
    st.enable_full_relativity = full_relativity
    grid = np.asarray(model.spectrum_frequency_grid)
    em = res["output_energies"] >= 0
    if fused:
        he = np.histogram(res["output_nus"][em], weights=res["output_energies"][em], bins=grid)[0]
        hr = np.histogram(res["output_nus"][~em], weights=-res["output_energies"][~em], bins=grid)[0]
        st.fused_packet_sums = mc.FusedPacketSums(he, hr, np.zeros(4), 0.0, np.inf)
    return model, res, st, grid


def reference_histograms(st, grid):
    """spectrum/base.py:139-159, literally"""
    em = np.histogram(np.asarray(sp._value(st.emitted_packet_nu)), weights=np.asarray(sp._value(st.emitted_packet_luminosity)), bins=grid)[0]
    re = np.histogram(np.asarray(sp._value(st.reabsorbed_packet_nu)), weights=np.asarray(sp._value(st.reabsorbed_packet_luminosity)), bins=grid)[0]
    return em, re


def test_real_packet_spectra_from_the_fused_histograms_equal_the_reference_formula(oracle):
    model, res, st, grid = state_from_oracle(oracle)
    solver = sp.SpectrumSolverB200(None, model.spectrum_frequency_grid, ns(compute="GPU"))
    solver.setup_optional_spectra(st)
    em, re = reference_histograms(st, grid)
    got_em, got_re = np.asarray(sp._value(solver.montecarlo_emitted_luminosity)), np.asarray(sp._value(solver.montecarlo_reabsorbed_luminosity))
    assert got_em.shape == (len(grid) - 1,) and em.sum() > 0 and re.sum() > 0
    # numpy's weighted histogram differences a cumulative sum over ALL packets: each bin carries ~eps x the total, and dividing by t
    # before or after moves that rounding -- the reference's own formula is only this exact
    np.testing.assert_allclose(got_em, em, rtol=0, atol=1e-14 * em.sum())
    np.testing.assert_allclose(got_re, re, rtol=0, atol=1e-14 * re.sum())
    assert np.array_equal(got_em == 0, em == 0) and np.array_equal(got_re == 0, re == 0)
    s = solver.spectrum_real_packets
    assert np.array_equal(np.asarray(sp._value(s._frequency)), grid) and np.array_equal(np.asarray(sp._value(s.luminosity)), got_em)
    assert np.array_equal(np.asarray(sp._value(solver.spectrum_real_packets_reabsorbed.luminosity)), got_re)
    # the per-packet arrays are not needed for that: drop them and ask again
    st.packet_collection.output_nus = st.packet_collection.output_energies = None
    assert np.array_equal(np.asarray(sp._value(solver.montecarlo_emitted_luminosity)), got_em)


def test_a_state_without_fused_sums_takes_the_reference_formula(oracle):
    model, res, st, grid = state_from_oracle(oracle, fused=False)
    solver = sp.SpectrumSolverB200(st, model.spectrum_frequency_grid, ns(compute="GPU"))
    em, re = reference_histograms(st, grid)
    assert np.array_equal(np.asarray(sp._value(solver.montecarlo_emitted_luminosity)), em)
    assert np.array_equal(np.asarray(sp._value(solver.montecarlo_reabsorbed_luminosity)), re)
    # fused sums on another grid than the solver's are not used either
    st.fused_packet_sums = mc.FusedPacketSums(np.ones(5), np.ones(5), np.zeros(4), 0.0, np.inf)
    assert np.array_equal(np.asarray(sp._value(solver.montecarlo_emitted_luminosity)), em)


def test_virtual_spectrum_and_solve(oracle):
    model, res, st, grid = state_from_oracle(oracle, vpackets=3)
    solver = sp.SpectrumSolverB200(None, model.spectrum_frequency_grid, ns(compute="GPU"))
    solver.transport_state = st
    with pytest.warns(UserWarning, match="spectrum_virtual_packets is zero"):
        assert np.all(np.asarray(sp._value(solver.spectrum_virtual_packets.luminosity)) == 0)
    vhist = res["vhist"]
    assert vhist.shape == (len(grid),) and vhist.sum() > 0
    solver.setup_optional_spectra(st, vhist)
    want = vhist[:-1] / st.packet_collection.time_of_simulation  # spectrum/base.py:162-166
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        real, virtual, integrated = solver.solve(st)
    assert np.array_equal(np.asarray(sp._value(virtual.luminosity)), want)
    assert np.array_equal(np.asarray(sp._value(real.luminosity)), np.asarray(sp._value(solver.montecarlo_emitted_luminosity)))
    assert integrated is None and any("FormalIntegrator is not yet available" in str(x.message) for x in w)


def test_integrated_spectrum_calls_the_integrator_like_the_reference(oracle):
    model, res, st, grid = state_from_oracle(oracle)
    calls = []

    class Integrator:
        def solve(self, *args):
            calls.append(args)
            return "spectrum"

    solver = sp.SpectrumSolverB200(None, model.spectrum_frequency_grid, ns(compute="GPU"))
    plasma = ns(atomic_data="atomic", electron_densities="ne")
    solver.setup_optional_spectra(st, None, Integrator(), simulation_state="sim", transport="transport", plasma=plasma, opacity_state="opacity",
                                  macro_atom_state="macro")
    assert solver.spectrum_integrated == "spectrum" and solver.spectrum_integrated == "spectrum" and len(calls) == 1  # cached
    nu, *rest = calls[0]
    assert np.array_equal(np.asarray(nu), grid[:-1]) and rest == ["sim", "transport", "opacity", "atomic", "ne", "macro"]

    class Refusing:
        def solve(self, *args):
            raise IntegrationError("continuum")

    solver.setup_optional_spectra(st, None, Refusing(), plasma=plasma)
    with pytest.warns(UserWarning, match="RETURNS AN EMPTY SPECTRUM"):
        empty = solver.spectrum_integrated
    assert np.isnan(np.asarray(sp._value(empty.luminosity))).all() and np.shape(sp._value(empty._frequency)) == (2,)
    # full relativity: the reference refuses to hand out an integrator at all (spectrum/base.py:130-136)
    model, res, st, grid = state_from_oracle(oracle, full_relativity=True)
    solver.setup_optional_spectra(st, None, Integrator(), plasma=plasma)
    with pytest.raises(NotImplementedError, match="full relativity"):
        solver.spectrum_integrated


def test_tardis_spectrum_fields():
    edges = np.linspace(1.0e14, 2.0e14, 11)
    lum = np.arange(10, dtype=np.float64) + 1.0
    s = sp.TARDISSpectrumB200(edges, lum)
    assert np.array_equal(np.asarray(sp._value(s.frequency)), edges[:-1]) and float(sp._value(s.delta_frequency)) == edges[1] - edges[0]
    np.testing.assert_allclose(np.asarray(sp._value(s.luminosity_density_nu)), lum / 1.0e13, rtol=1e-15)
    np.testing.assert_allclose(np.asarray(sp._value(s.wavelength)), 2.99792458e18 / edges[:-1], rtol=1e-15)
    # L_lambda dlambda == L_nu dnu: f_nu_to_f_lambda is L_nu nu^2 / c (spectrum.py:96-97)
    np.testing.assert_allclose(np.asarray(sp._value(s.luminosity_density_lambda)), (lum / 1.0e13) * edges[:-1] ** 2 / 2.99792458e18, rtol=1e-15)
    with pytest.raises(ValueError, match="not compatible"):
        sp.TARDISSpectrumB200(edges, lum[:-1])


def test_from_config_builds_the_transport_solver_s_grid():
    cfg = ns(spectrum=ns(start=1.0e14, stop=3.0e15, num=100, integrated=ns(compute="GPU", points=1000, interpolate_shells=0)))
    solver = sp.SpectrumSolverB200.from_config(cfg)
    grid = np.asarray(sp._value(solver.spectrum_frequency_grid))
    assert grid.shape == (101,) and grid[0] == 3.0e15 and grid[-1] == 1.0e14
    assert solver.integrator_settings.points == 1000 and solver.transport_state is None
    assert np.asarray(sp._value(solver._montecarlo_virtual_luminosity)).shape == (101,)
