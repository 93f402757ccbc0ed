"""CPU-only tests of the host-side mirror of the reference interface (tardis_b200/montecarlo.py):
argument marshalling, configuration_initialize, state properties -- with a recording fake engine
(no compute happens on the CPU; the real engine needs a GPU)."""
from types import SimpleNamespace

import numpy as np
import pytest

from tardis_b200 import montecarlo as mc
from tardis_b200 import synthetic as syn


class FakeEngine:
    def __init__(self):
        self.model_kwargs = None
        self.run_args = None

    def set_model(self, **kw):
        self.model_kwargs = kw

    def run(self, r, nu, mu, e, seeds, **opts):
        self.run_args = (r, nu, mu, e, seeds, opts)
        n = len(nu)
        L, S = self.model_kwargs["tau_sobolev"].shape
        G = len(self.model_kwargs["spectrum_frequency_grid"])
        res = dict(j=np.arange(S, dtype=float), nu_bar=np.ones(S), j_blue=np.zeros((L, S)), edotlu=np.zeros((L, S)),
                   vhist=np.zeros(G), counters={}, vlog_count=0)
        if "buffers" in opts and opts["buffers"]:
            opts["buffers"]["output_nus"][:] = 2.0 * np.asarray(nu)
            opts["buffers"]["output_energies"][:] = np.where(np.arange(n) % 2, 1.0, -1.0) * np.asarray(e)
            res["output_nus"], res["output_energies"] = opts["buffers"]["output_nus"], opts["buffers"]["output_energies"]
        if opts.get("track_last_interaction"):
            for k in mc.LastInteractionTrackers.INT_COLUMNS:
                res[k] = np.full(n, 2 if k == "last_interaction_type" else 0, dtype=np.int64)
            for k in mc.LastInteractionTrackers.FLOAT_COLUMNS:
                res[k] = np.zeros(n)
        return res


def _setup(mode="macroatom"):
    model = syn.make_model(6, 300, mode, mu_tau=-3.0, seed=1)
    packets = syn.make_packets(50, model.r_inner[0])
    geo = mc.HomologousGeometry(model.r_inner, model.r_outer, model.v_inner, model.v_outer, model.time_explosion)
    opa = mc.OpacityState.from_model(model)
    pc = mc.PacketCollection(packets.initial_radii, packets.initial_nus, packets.initial_mus, packets.initial_energies,
                             packets.packet_seeds, packets.radiation_field_luminosity)
    return model, geo, opa, pc


def test_transport_call_marshals_reference_arguments():
    model, geo, opa, pc = _setup()
    cfg = mc.MonteCarloConfiguration()
    cfg.LINE_INTERACTION_TYPE = 2
    cfg.NUMBER_OF_VPACKETS = 3
    cfg.ENABLE_FULL_RELATIVITY = True
    cfg.VPACKET_SPAWN_START_FREQUENCY, cfg.VPACKET_SPAWN_END_FREQUENCY = 1e14, 2e15
    eng = FakeEngine()
    trackers = mc.generate_tracker_last_interaction_list(len(pc.initial_nus))
    vhist, vtracker, bulk, line = mc.montecarlo_transport_with_vpackets(
        pc, geo, model.time_explosion, opa, cfg, model.spectrum_frequency_grid, trackers, 3, False, None, engine=eng)
    kw = eng.model_kwargs
    assert kw["line_interaction_type"] == 2 and kw["number_of_vpackets"] == 3 and kw["enable_full_relativity"] is True
    assert kw["vpacket_spawn_start_frequency"] == 1e14 and kw["vpacket_spawn_end_frequency"] == 2e15
    assert kw["tau_sobolev"] is model.tau_sobolev and kw["time_explosion"] == model.time_explosion
    assert kw["sigma_thomson"] == mc.SIGMA_THOMSON
    # outputs written in place, sign convention preserved
    assert np.array_equal(pc.output_nus, 2.0 * pc.initial_nus)
    assert (pc.output_energies[1::2] > 0).all() and (pc.output_energies[0::2] < 0).all()
    assert bulk.mean_intensity_total.shape == (6,) and line.mean_intensity_blueward.shape == (300, 6)
    assert vhist.shape == model.spectrum_frequency_grid.shape
    df = trackers.to_df()
    assert list(df.columns) == ["event_id", "last_interaction_type", "status", "radius", "shell_id", "before_nu", "before_mu",
                                "before_energy", "after_nu", "after_mu", "after_energy", "line_absorb_id", "line_emit_id"]
    assert (df["last_interaction_type"] == "LINE").all()


def test_opacity_state_shell_slice_is_a_view():
    model, geo, opa, pc = _setup()
    sl = opa[1:4]
    assert sl.tau_sobolev.shape == (300, 3) and not sl.tau_sobolev.flags["C_CONTIGUOUS"]
    assert sl.tau_sobolev.base is not None
    assert sl.transition_probabilities.shape[1] == 3
    assert sl.line_list_nu is opa.line_list_nu


def test_configuration_initialize_swaps_spawn_range():
    cfg = mc.MonteCarloConfiguration()
    transport = SimpleNamespace(line_interaction_type="downbranch", enable_full_relativity=False,
                                packet_source=SimpleNamespace(base_seed=7),
                                vpacket_spawn_range=SimpleNamespace(start=3e15, end=1e14), enable_vpacket_tracking=True)
    mc.configuration_initialize(cfg, transport, 5)
    assert cfg.LINE_INTERACTION_TYPE == 1 and cfg.NUMBER_OF_VPACKETS == 5 and cfg.TEMPORARY_V_PACKET_BINS == 5
    assert cfg.VPACKET_SPAWN_START_FREQUENCY == 1e14 and cfg.VPACKET_SPAWN_END_FREQUENCY == 3e15
    assert cfg.MONTECARLO_SEED == 7 and cfg.ENABLE_VPACKET_TRACKING is True
    transport.line_interaction_type = "bogus"
    with pytest.raises(ValueError):
        mc.configuration_initialize(cfg, transport, 0)


def _fake_config(**over):
    c = SimpleNamespace(
        plasma=SimpleNamespace(disable_electron_scattering=False, disable_line_scattering=False, w_epsilon=1e-10,
                               line_interaction_type="macroatom"),
        spectrum=SimpleNamespace(start=1e15, stop=1e14, num=100, method="real", integrated=SimpleNamespace(compute="Automatic"),
                                 virtual=SimpleNamespace(virtual_packet_logging=False)),
        montecarlo=SimpleNamespace(virtual_spectrum_spawn_range=SimpleNamespace(start=3e15, end=1e13),
                                   enable_full_relativity=False, debug_packets=False, logger_buffer=1, nthreads=1,
                                   tracking=SimpleNamespace(track_rpacket=False, initial_array_length=10)))
    for k, v in over.items():
        obj = c
        parts = k.split(".")
        for part in parts[:-1]:
            obj = getattr(obj, part)
        setattr(obj, parts[-1], v)
    return c


def test_solver_from_config_and_run(monkeypatch):
    model, geo, opa, pc = _setup()
    eng = FakeEngine()
    monkeypatch.setattr(mc, "get_engine", lambda device=0: eng)
    source = SimpleNamespace(base_seed=23111963, create_packets=lambda n, seed_offset=0: pc)
    solver = mc.MCTransportSolverB200.from_config(_fake_config(), source)
    grid = np.asarray(getattr(solver.spectrum_frequency_grid, "value", solver.spectrum_frequency_grid))
    assert len(grid) == 101 and grid[0] == 1e14 and grid[-1] == 1e15
    sim_state = SimpleNamespace(geometry=SimpleNamespace(to_numba=lambda: geo, v_inner_boundary_idx=0, v_outer_boundary_idx=6),
                                time_explosion=model.time_explosion)
    opacity_host = SimpleNamespace(to_numba=lambda macro_atom_state, lit: opa)
    plasma = SimpleNamespace(continuum_interaction_species=SimpleNamespace(empty=True))
    state = solver.initialize_transport_state(sim_state, opacity_host, None, plasma, 50, no_of_virtual_packets=2, iteration=3)
    assert solver.montecarlo_configuration.NUMBER_OF_VPACKETS == 2
    assert solver.montecarlo_configuration.LINE_INTERACTION_TYPE == 2
    vhist = solver.run(state, show_progress_bars=False)
    assert solver.transport_state is state
    assert state.estimators_bulk is not None and state.j_estimator.shape == (6,)
    assert state.j_blue_estimator.shape == (300, 6)
    assert state.tracker_last_interaction_df.shape[0] == 50
    assert np.array_equal(np.asarray(getattr(state.output_nu, "value", state.output_nu)), pc.output_nus)
    assert state.emitted_packet_mask.sum() == 25
    assert eng.model_kwargs["number_of_vpackets"] == 2
    assert vhist.shape == (101,)


def test_solver_disable_electron_scattering_follows_reference_effective_behaviour():
    source = SimpleNamespace(base_seed=1)
    s1 = mc.MCTransportSolverB200.from_config(_fake_config(**{"plasma.disable_electron_scattering": True}), source)
    assert s1.sigma_thomson == mc.SIGMA_THOMSON  # the reference's monkey patch never reaches its compiled loop
    s2 = mc.MCTransportSolverB200.from_config(_fake_config(**{"plasma.disable_electron_scattering": True}), source,
                                              honor_disable_electron_scattering=True)
    assert s2.sigma_thomson == 1e-200
    with pytest.raises(ValueError):
        mc.MCTransportSolverB200.from_config(_fake_config(**{"spectrum.integrated.compute": "TPU"}), source)
    with pytest.raises(NotImplementedError):
        mc.MCTransportSolverB200.from_config(_fake_config(**{"montecarlo.tracking.track_rpacket": True}), source)
