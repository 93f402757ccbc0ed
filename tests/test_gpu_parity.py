"""GPU parity tests: the CUDA engine (through the C-ABI) against (a) golden vectors from the
reference Numba path and (b) the CPU oracle on seeded synthetic inputs.  Integer work
(trajectories, event counts, work counters, last-interaction ids) must be identical; floats agree
to the tolerance in golden_util (the reference's own regression tolerance is 1e-12...1e-13)."""
import numpy as np
import pytest

from golden_util import CASES, assert_close, compare_to_golden, load_case, make_golden, oracle_kwargs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module", params=["scan", "jump"])
def engine(request):
    """Both transport algorithms must produce the reference's results: `scan` streams the line list,
    `jump` searches the tau prefix table and range-updates the line estimators."""
    from tardis_b200.engine import Engine

    eng = Engine(0)
    eng.set_option("algorithm", {"scan": 0, "jump": 1}[request.param])
    yield eng
    eng.close()


def engine_config(rk, sig):
    cfg = dict(
        enable_full_relativity=rk.get("enable_full_relativity", False),
        disable_line_scattering=rk.get("disable_line_scattering", False),
        number_of_vpackets=rk.get("number_of_vpackets", 0),
        vpacket_spawn_start_frequency=rk.get("spawn_start", 0.0),
        vpacket_spawn_end_frequency=rk.get("spawn_end", 1e200),
        survival_probability=rk.get("survival_probability", 0.0),
    )
    if sig is not None:
        cfg["sigma_thomson"] = sig
    return cfg


@pytest.mark.parametrize("name", CASES)
def test_engine_matches_reference_golden(engine, name):
    model, packets, rk, sig, g = load_case(name)
    engine.set_model_from(model, **engine_config(rk, sig))
    res = engine.run_packets(packets, track_last_interaction=True, n_tracked_packets=make_golden.N_TRACKED,
                             max_events_per_packet=4096)
    compare_to_golden(res, g, make_golden.N_TRACKED)


@pytest.mark.parametrize("name", CASES)
def test_engine_matches_oracle_counters(engine, oracle, name):
    """Work counters are functions of the seeds only: the engine must report the oracle's numbers."""
    model, packets, rk, sig, g = load_case(name)
    ref = oracle.run_oracle(model, packets, **oracle_kwargs(rk, sig))
    engine.set_model_from(model, **engine_config(rk, sig))
    res = engine.run_packets(packets)
    assert res["counters"] == ref["counters"]


@pytest.mark.parametrize("mode,n_lines,mu_tau", [("scatter", 20000, -6.0), ("macroatom", 8000, -4.5), ("downbranch", 8000, -4.5)])
def test_engine_vs_oracle_larger(engine, oracle, mode, n_lines, mu_tau):
    from tardis_b200 import synthetic as syn

    model = syn.make_model(20, n_lines, mode, mu_tau=mu_tau, seed=777)
    packets = syn.make_packets(20000, model.r_inner[0], base_seed=99)
    ref = oracle.run_oracle(model, packets, nthreads=8, n_tracked_packets=500, max_events_per_packet=4096)
    engine.set_model_from(model)
    res = engine.run_packets(packets, track_last_interaction=True, n_tracked_packets=500, max_events_per_packet=4096)
    assert res["counters"] == ref["counters"]
    assert_close(res["output_nus"], ref["output_nus"], 1e-11, "output_nus")
    assert_close(res["output_energies"], ref["output_energies"], 1e-11, "output_energies")
    for k in ("j", "nu_bar", "j_blue", "edotlu"):
        assert_close(res[k], ref[k], 1e-10, k)
    for k in ("last_interaction_type", "last_event_id", "last_shell_id", "last_line_absorb_id", "last_line_emit_id"):
        assert np.array_equal(res[k], ref[k]), k
    assert np.array_equal(res["event_counts"], ref["event_counts"])
    a, b = np.concatenate(res["events"]), np.concatenate(ref["events"])
    for k in ("packet_id", "interaction_type", "status", "before_shell_id", "after_shell_id", "line_absorb_id", "line_emit_id"):
        assert np.array_equal(a[k], b[k]), k


def test_engine_rng_long_stream(engine, oracle):
    """A very optically thick single shell forces hundreds of draws per packet: exercises all three
    tiers of the lazy MT19937 (cursor-only, ring read-back, in-place recurrence)."""
    from tardis_b200 import synthetic as syn

    model = syn.make_model(3, 2000, "scatter", mu_tau=-1.0, seed=5)
    model.electron_density[:] *= 300.0
    packets = syn.make_packets(2000, model.r_inner[0], base_seed=7)
    ref = oracle.run_oracle(model, packets)
    engine.set_model_from(model)
    res = engine.run_packets(packets)
    assert ref["counters"]["n_rng_draws"] / len(packets) > 150
    draws = ref["counters"]["n_rng_draws"]
    assert res["counters"] == ref["counters"], (res["counters"], ref["counters"], draws)
    assert_close(res["output_nus"], ref["output_nus"], 1e-10, "output_nus")
    assert_close(res["output_energies"], ref["output_energies"], 1e-10, "output_energies")


def test_empty_and_single_packet(engine, oracle):
    from tardis_b200 import synthetic as syn

    model = syn.make_model(5, 500, "scatter", mu_tau=-3.0, seed=11)
    engine.set_model_from(model)
    p0 = syn.make_packets(1, model.r_inner[0])
    res = engine.run_packets(p0.slice(0, 0))
    assert res["output_nus"].shape == (0,) and res["j"].sum() == 0.0
    res = engine.run_packets(p0)
    ref = oracle.run_oracle(model, p0)
    assert_close(res["output_nus"], ref["output_nus"], 1e-12, "nu")
    assert res["counters"] == ref["counters"]


def test_ragged_warp_fill(engine, oracle):
    """Packet counts that are not multiples of the warp / grid size."""
    from tardis_b200 import synthetic as syn

    model = syn.make_model(6, 1500, "scatter", mu_tau=-3.5, seed=12)
    engine.set_model_from(model)
    for n in (31, 33, 1000, 37889):
        pk = syn.make_packets(n, model.r_inner[0], base_seed=n)
        res = engine.run_packets(pk)
        ref = oracle.run_oracle(model, pk, nthreads=4)
        assert res["counters"] == ref["counters"]
        assert_close(res["output_energies"], ref["output_energies"], 1e-11, "energies")


def test_sliced_noncontiguous_tau(engine, oracle):
    """The reference hands tau_sobolev[:, i:j] as a non-contiguous view (modes/classic/solver.py:132-134)."""
    from tardis_b200 import synthetic as syn

    full = syn.make_model(12, 2000, "macroatom", mu_tau=-4.0, seed=13)
    sl = slice(2, 9)
    view = syn.Model(
        r_inner=full.r_inner[sl].copy(), r_outer=full.r_outer[sl].copy(), v_inner=full.v_inner[sl].copy(),
        v_outer=full.v_outer[sl].copy(), time_explosion=full.time_explosion,
        electron_density=full.electron_density[sl].copy(), t_electrons=full.t_electrons[sl].copy(),
        line_list_nu=full.line_list_nu, tau_sobolev=full.tau_sobolev[:, sl],
        macro=syn.MacroAtomTables(full.macro.transition_probabilities[:, sl], full.macro.line2macro_level_upper,
                                  full.macro.macro_block_edge_index, full.macro.transition_type,
                                  full.macro.destination_level_id, full.macro.transition_line_id),
        spectrum_frequency_grid=full.spectrum_frequency_grid, line_interaction_type="macroatom")
    assert not view.tau_sobolev.flags["C_CONTIGUOUS"]
    pk = syn.make_packets(3000, view.r_inner[0], base_seed=21)
    engine.set_model_from(view)
    res = engine.run_packets(pk)
    ref = oracle.run_oracle(view, pk, nthreads=4)
    assert res["counters"] == ref["counters"]
    assert_close(res["j_blue"], ref["j_blue"], 1e-10, "j_blue")


def test_error_conditions(engine):
    from tardis_b200 import synthetic as syn
    from tardis_b200.engine import EngineError, MacroAtomError

    model = syn.make_model(4, 300, "macroatom", mu_tau=-2.0, seed=14)
    bad = syn.make_model(4, 300, "macroatom", mu_tau=-2.0, seed=14)
    bad.macro.transition_probabilities[:] = 0.0  # every activated block runs out -> MacroAtomError in the reference
    engine.set_model_from(bad)
    with pytest.raises(MacroAtomError):
        engine.run_packets(syn.make_packets(500, bad.r_inner[0]))
    unsorted = syn.make_model(4, 300, "scatter", seed=15)
    unsorted.line_list_nu[10], unsorted.line_list_nu[200] = unsorted.line_list_nu[200], unsorted.line_list_nu[10]
    with pytest.raises(EngineError):
        engine.set_model_from(unsorted)
    engine.set_model_from(model)  # engine stays usable
    engine.run_packets(syn.make_packets(100, model.r_inner[0]))
