"""Run the UNMODIFIED reference Monte Carlo loop on a synthetic model.

TEST INFRASTRUCTURE ONLY (see oracle/reference_loader.py).  Builds the
reference's own jitclass carriers from a `tardis_b200.synthetic.Model` and calls
`montecarlo_transport_with_vpackets`
(tardis/transport/montecarlo/modes/montecarlo_transport.py:239) with the
reference's `packet_propagation`
(tardis/transport/montecarlo/modes/classic/packet_propagation.py:53).
"""
from __future__ import annotations

import numpy as np

from . import reference_loader

LINE_INTERACTION = {"scatter": 0, "downbranch": 1, "macroatom": 2}


def set_sigma_thomson(value: float) -> None:
    """Give the reference a different Thomson cross-section (1e-200 = "electron
    scattering disabled", modes/classic/solver.py:291-300).

    NOTE on the reference's own behaviour: `from_config` assigns
    `constants.SIGMA_THOMSON = 1e-200`, but every consumer on the hot path did
    `from ...configuration.constants import SIGMA_THOMSON` at import time
    (opacities/opacities.py:10, packets/virtual_packet.py:22), so in a real
    run that assignment never reaches the compiled loop.  To exercise the
    *intended* behaviour we patch the importing modules' own globals.  Must be
    called BEFORE the first JIT compile in the process: module constants are
    frozen at compile time."""
    reference_loader.load()
    import tardis.opacities.opacities as opacities
    import tardis.transport.montecarlo.configuration.constants as constants
    import tardis.transport.montecarlo.packets.virtual_packet as virtual_packet

    constants.SIGMA_THOMSON = value
    opacities.SIGMA_THOMSON = value
    virtual_packet.SIGMA_THOMSON = value


def build_reference_objects(model, packets, *, number_of_vpackets=0, enable_full_relativity=False,
                            disable_line_scattering=False, survival_probability=0.0,
                            spawn_start=0.0, spawn_end=1e200):
    R = reference_loader.load()
    geometry = R.NumbaHomologousRadial1DGeometry(
        model.r_inner, model.r_outer, model.v_inner, model.v_outer, model.time_explosion
    )
    m = model.macro
    z1 = np.zeros(0, dtype=np.float64)
    z2 = np.zeros((0, 0), dtype=np.float64)
    zi = np.zeros(0, dtype=np.int64)
    opacity = R.OpacityStateNumba(
        model.electron_density, model.t_electrons, model.line_list_nu, model.tau_sobolev,
        m.transition_probabilities, m.line2macro_level_upper, m.macro_block_edge_index,
        m.transition_type, m.destination_level_id, m.transition_line_id,
        z1, z2, z1, z1, zi, z2, z1, z1, z1, z2, zi, np.int64(-1),
    )
    cfg = R.MonteCarloConfiguration()
    cfg.ENABLE_FULL_RELATIVITY = enable_full_relativity
    cfg.NUMBER_OF_VPACKETS = number_of_vpackets
    cfg.TEMPORARY_V_PACKET_BINS = number_of_vpackets
    cfg.LINE_INTERACTION_TYPE = LINE_INTERACTION[model.line_interaction_type]
    cfg.DISABLE_LINE_SCATTERING = disable_line_scattering
    cfg.SURVIVAL_PROBABILITY = survival_probability
    cfg.VPACKET_SPAWN_START_FREQUENCY = spawn_start
    cfg.VPACKET_SPAWN_END_FREQUENCY = spawn_end
    pc = R.PacketCollection(
        packets.initial_radii.copy(), packets.initial_nus.copy(), packets.initial_mus.copy(),
        packets.initial_energies.copy(), packets.packet_seeds.copy(), packets.radiation_field_luminosity,
    )
    return R, geometry, opacity, cfg, pc


def run_reference(model, packets, *, number_of_vpackets=0, enable_full_relativity=False,
                  disable_line_scattering=False, survival_probability=0.0,
                  spawn_start=0.0, spawn_end=1e200, track_full=False, nthreads=1, timings=None):
    """Returns a dict of numpy outputs of one reference MC iteration (`timings`, if a dict, receives the seconds of the
    reference call itself as "loop_s" and of its per-packet tracker construction as "tracker_setup_s")."""
    import time

    import numba

    R, geometry, opacity, cfg, pc = build_reference_objects(
        model, packets, number_of_vpackets=number_of_vpackets,
        enable_full_relativity=enable_full_relativity,
        disable_line_scattering=disable_line_scattering,
        survival_probability=survival_probability, spawn_start=spawn_start, spawn_end=spawn_end,
    )
    numba.set_num_threads(nthreads)
    n = len(packets)
    t0 = time.perf_counter()
    if track_full:
        trackers = R.generate_tracker_full_list(n, 10)
    else:
        trackers = R.generate_tracker_last_interaction_list(n)
    t1 = time.perf_counter()
    vhist, vtracker, bulk, line = R.montecarlo_transport_with_vpackets(
        pc, geometry, model.time_explosion, opacity, cfg, model.spectrum_frequency_grid,
        trackers, number_of_vpackets, False, R.packet_propagation,
    )
    if timings is not None:
        timings["tracker_setup_s"], timings["loop_s"] = t1 - t0, time.perf_counter() - t1
    out = dict(
        output_nus=np.asarray(pc.output_nus).copy(),
        output_energies=np.asarray(pc.output_energies).copy(),
        j=np.asarray(bulk.mean_intensity_total).copy(),
        nu_bar=np.asarray(bulk.mean_frequency).copy(),
        j_blue=np.asarray(line.mean_intensity_blueward).copy(),
        edotlu=np.asarray(line.energy_deposition_line_rate).copy(),
        vhist=np.asarray(vhist).copy(),
    )
    if track_full:
        df = R.trackers_full_to_df(trackers)
        out["events"] = df
    else:
        out["last_interaction_type"] = np.array([t.interaction_type for t in trackers])
        out["last_event_id"] = np.array([t.interactions_count for t in trackers])
        out["last_radius"] = np.array([t.radius for t in trackers])
        out["last_shell_id"] = np.array([t.shell_id for t in trackers])
        out["last_before_nu"] = np.array([t.before_nu for t in trackers])
        out["last_before_mu"] = np.array([t.before_mu for t in trackers])
        out["last_before_energy"] = np.array([t.before_energy for t in trackers])
        out["last_after_nu"] = np.array([t.after_nu for t in trackers])
        out["last_after_mu"] = np.array([t.after_mu for t in trackers])
        out["last_after_energy"] = np.array([t.after_energy for t in trackers])
        out["last_line_absorb_id"] = np.array([t.interaction_line_absorb_id for t in trackers])
        out["last_line_emit_id"] = np.array([t.interaction_line_emit_id for t in trackers])
        out["boundary_buffer"] = np.array([t._boundary_interactions_buffer for t in trackers])
    return out


def run_reference_iip(model, packets, *, disable_line_scattering=False, track_full=False, nthreads=1, timings=None):
    """One MC iteration of the reference's IIP (continuum) mode: `montecarlo_transport`
    (tardis/transport/montecarlo/modes/iip/montecarlo_transport.py:40-176) with its own IIP `packet_propagation`
    (modes/iip/packet_propagation.py:55-270).  `model.continuum` must be set (tardis_b200.synthetic.add_continuum)."""
    import numba

    R = reference_loader.load()
    import tardis.transport.montecarlo.configuration.montecarlo_globals as montecarlo_globals

    montecarlo_globals.CONTINUUM_PROCESSES_ENABLED = True  # modes/iip/solver.py:132 (frozen at first JIT compile)
    from tardis.opacities.opacity_state_numba_iip import OpacityStateNumbaIIP
    from tardis.transport.montecarlo.modes.iip.montecarlo_transport import montecarlo_transport

    c, m = model.continuum, model.macro
    geometry = R.NumbaHomologousRadial1DGeometry(model.r_inner, model.r_outer, model.v_inner, model.v_outer, model.time_explosion)
    opacity = OpacityStateNumbaIIP(
        model.electron_density, model.t_electrons, model.line_list_nu, model.tau_sobolev,
        m.transition_probabilities, m.line2macro_level_upper, m.macro_block_edge_index, m.transition_type,
        m.destination_level_id, m.transition_line_id, c.bf_threshold_list_nu, c.p_fb_deactivation,
        c.photo_ion_nu_threshold_mins, c.photo_ion_nu_threshold_maxs, c.photo_ion_block_references, c.chi_bf, c.x_sect,
        c.phot_nus, c.ff_opacity_factor, c.emissivities, c.photo_ion_activation_idx, np.int64(c.k_packet_idx),
        c.absorbing_markov_probabilities)
    cfg = R.MonteCarloConfiguration()
    cfg.ENABLE_FULL_RELATIVITY = True
    cfg.LINE_INTERACTION_TYPE = LINE_INTERACTION[model.line_interaction_type]
    cfg.DISABLE_LINE_SCATTERING = disable_line_scattering
    pc = R.PacketCollection(packets.initial_radii.copy(), packets.initial_nus.copy(), packets.initial_mus.copy(),
                            packets.initial_energies.copy(), packets.packet_seeds.copy(), packets.radiation_field_luminosity)
    numba.set_num_threads(nthreads)
    n = len(packets)
    import time

    t0 = time.perf_counter()
    trackers = R.generate_tracker_full_list(n, 10) if track_full else R.generate_tracker_last_interaction_list(n)
    t1 = time.perf_counter()
    bulk, line, cont = montecarlo_transport(pc, geometry, model.time_explosion, opacity, cfg,
                                            (len(c.bf_threshold_list_nu), model.n_shells), trackers, False)
    if timings is not None:
        timings["tracker_setup_s"], timings["loop_s"] = t1 - t0, time.perf_counter() - t1
    out = dict(
        output_nus=np.asarray(pc.output_nus).copy(), output_energies=np.asarray(pc.output_energies).copy(),
        j=np.asarray(bulk.mean_intensity_total).copy(), nu_bar=np.asarray(bulk.mean_frequency).copy(),
        j_blue=np.asarray(line.mean_intensity_blueward).copy(), edotlu=np.asarray(line.energy_deposition_line_rate).copy(),
        vhist=np.zeros_like(model.spectrum_frequency_grid),
    )
    for k in ("photo_ion_estimator", "stim_recomb_estimator", "bf_heating_estimator", "stim_recomb_cooling_estimator",
              "ff_heating_estimator", "photo_ion_estimator_statistics"):
        out[k] = np.asarray(getattr(cont, k)).copy()
    if track_full:
        out["events"] = R.trackers_full_to_df(trackers)
    else:
        out["last_interaction_type"] = np.array([t.interaction_type for t in trackers])
        out["last_event_id"] = np.array([t.interactions_count for t in trackers])
        out["last_radius"] = np.array([t.radius for t in trackers])
        out["last_shell_id"] = np.array([t.shell_id for t in trackers])
        out["last_before_nu"] = np.array([t.before_nu for t in trackers])
        out["last_before_mu"] = np.array([t.before_mu for t in trackers])
        out["last_before_energy"] = np.array([t.before_energy for t in trackers])
        out["last_after_nu"] = np.array([t.after_nu for t in trackers])
        out["last_after_mu"] = np.array([t.after_mu for t in trackers])
        out["last_after_energy"] = np.array([t.after_energy for t in trackers])
        out["last_line_absorb_id"] = np.array([t.interaction_line_absorb_id for t in trackers])
        out["last_line_emit_id"] = np.array([t.interaction_line_emit_id for t in trackers])
    return out


def run_reference_packet_source(no_of_packets, base_seed, seed_offset, radius, temperature, time_explosion=None):
    """The unmodified `BlackBodySimpleSource.create_packets` (packet_source/base.py:195-253, black_body.py:122-220).

    Two of its imports do not exist in this container and are stood in for: `numexpr` (its one use,
    `ne.evaluate("-log(xis_prod)/l")`, is evaluated with numpy -- so the goldens carry numpy's `log`, which may differ
    from numexpr's by an ulp) and `tardis.io.hdf_writer_mixin` (an empty mixin class).  The stand-in Quantity gets
    `__array_ufunc__ = None` so that `ndarray * Quantity` defers to the Quantity, as astropy's does."""
    import sys
    import types

    from oracle import reference_loader

    reference_loader.load()
    reference_loader._Q.__array_ufunc__ = None
    if "numexpr" not in sys.modules:
        ne = types.ModuleType("numexpr")

        def evaluate(expr, local_dict=None):
            frame = sys._getframe(1)
            env = dict(frame.f_globals)
            env.update(frame.f_locals)
            env.update(local_dict or {})
            env["log"] = np.log
            return eval(expr, {"__builtins__": {}}, env)  # noqa: S307 -- the expression is the reference's own literal

        ne.evaluate = evaluate
        sys.modules["numexpr"] = ne
    if "tardis.io.hdf_writer_mixin" not in sys.modules:
        hm = types.ModuleType("tardis.io.hdf_writer_mixin")
        hm.HDFWriterMixin = type("HDFWriterMixin", (), {})
        sys.modules["tardis.io.hdf_writer_mixin"] = hm
    if "tardis.transport.montecarlo.packet_source" not in sys.modules:  # skip the package __init__ (it imports the gamma-ray sources)
        import os

        pkg = types.ModuleType("tardis.transport.montecarlo.packet_source")
        pkg.__path__ = [os.path.join(reference_loader.REF, "tardis", "transport", "montecarlo", "packet_source")]
        sys.modules["tardis.transport.montecarlo.packet_source"] = pkg
    from tardis.transport.montecarlo.packet_source.black_body import BlackBodySimpleSource

    if time_explosion is not None:  # BlackBodySimpleSourceRelativistic (black_body_relativistic.py), unmodified
        from tardis.transport.montecarlo.packet_source.black_body_relativistic import BlackBodySimpleSourceRelativistic

        src = BlackBodySimpleSourceRelativistic(time_explosion=reference_loader._Q(float(time_explosion)),
                                                radius=reference_loader._Q(float(radius)), temperature=reference_loader._Q(float(temperature)),
                                                base_seed=int(base_seed))
    else:
        src = BlackBodySimpleSource(radius=reference_loader._Q(float(radius)), temperature=reference_loader._Q(float(temperature)),
                                    base_seed=int(base_seed))
    pc = src.create_packets(int(no_of_packets), seed_offset=int(seed_offset))
    return dict(initial_radii=np.asarray(pc.initial_radii, dtype=np.float64).copy(), initial_nus=np.asarray(pc.initial_nus, dtype=np.float64).copy(),
                initial_mus=np.asarray(pc.initial_mus, dtype=np.float64).copy(),
                initial_energies=np.asarray(pc.initial_energies, dtype=np.float64).copy(),
                packet_seeds=np.asarray(pc.packet_seeds, dtype=np.int64).copy(),
                radiation_field_luminosity=float(pc.radiation_field_luminosity))


def run_reference_radfield(j, nu_bar, j_blue, time_explosion, time_of_simulation, volume, line_list_nu, w_epsilon=1e-10):
    """The unmodified `MCRadiationFieldPropertiesSolver.solve` (transport/montecarlo/estimators/mc_rad_field_solver.py:37-144,
    detailed_optical_window=False) with the unmodified `DilutePlanckianRadiationField` (plasma/radiation_field/planck_rad_field.py).

    One import of theirs does not exist in this container: `tardis.util.base` (it pulls astropy, radioactivedecay's data,
    pandas tables ...).  It is stood in for by a module holding only `intensity_black_body` (util/base.py:279-302) with its
    numexpr expression evaluated by numpy -- so the goldens carry numpy's `exp` / `**`, which may differ from numexpr's by an ulp."""
    import sys
    import types

    import numpy as np  # noqa: F811

    R = reference_loader.load()  # noqa: F841
    Q = reference_loader._Q
    Q.__array_ufunc__ = None
    if "tardis.util.base" not in sys.modules:
        ub = types.ModuleType("tardis.util.base")
        k_b, h, c = 1.3806488e-16, 6.62606957e-27, 2.99792458e10

        def intensity_black_body(nu, temperature):
            temperature = temperature.value if isinstance(temperature, Q) else temperature
            beta_rad = 1 / (k_b * temperature)
            coefficient = 2 * h / c**2
            return coefficient * nu**3 / (np.exp(h * nu * beta_rad) - 1)

        ub.intensity_black_body = intensity_black_body
        sys.modules["tardis.util.base"] = ub
    import tardis.constants as const

    if not hasattr(const, "sigma_sb"):
        const.sigma_sb = Q(5.670373e-5)
    import tardis.plasma.radiation_field as rf_pkg
    from tardis.plasma.radiation_field.planck_rad_field import DilutePlanckianRadiationField

    rf_pkg.DilutePlanckianRadiationField = DilutePlanckianRadiationField
    from tardis.transport.montecarlo.estimators.estimators_bulk import EstimatorsBulk
    from tardis.transport.montecarlo.estimators.estimators_line import EstimatorsLine
    from tardis.transport.montecarlo.estimators.mc_rad_field_solver import MCRadiationFieldPropertiesSolver

    bulk = EstimatorsBulk(np.array(j, dtype=np.float64), np.array(nu_bar, dtype=np.float64))
    line = EstimatorsLine(np.array(j_blue, dtype=np.float64), np.zeros_like(j_blue, dtype=np.float64))
    out = MCRadiationFieldPropertiesSolver(w_epsilon).solve(bulk, line, Q(float(time_explosion)), Q(float(time_of_simulation)),
                                                           np.asarray(volume, dtype=np.float64), np.asarray(line_list_nu, dtype=np.float64))
    st = out.dilute_blackbody_radiationfield_state
    t = st.temperature
    return dict(t_radiative=np.asarray(t.value if isinstance(t, Q) else t, dtype=np.float64),
                dilution_factor=np.asarray(st.dilution_factor, dtype=np.float64), j_blues=np.asarray(out.j_blues, dtype=np.float64))


def run_reference_opacity(atomic, plasma, nlte=False):
    """tau_Sobolev, beta_Sobolev and the raw macro-atom probabilities from the UNMODIFIED reference functions
    `calculate_sobolev_line_opacity`, `numba_calculate_beta_sobolev` (opacities/tau_sobolev.py:21-88) and
    `probability_emission_down / _internal_down / _internal_up` (opacities/macro_atom/macroatom_line_transitions.py).

    Two steps of the chain live in modules that cannot be imported here (plasma/properties/radiative_properties.py and
    opacities/macro_atom/macroatom_solver.py pull the atomic-data / HDF stack): the stimulated-emission factor is taken from
    oracle/opacity_oracle.py (a restatement), and the normalisation is the reference's own pandas expression
    (`df.div(df.groupby("source").transform("sum"))`, NaN -> 0; macroatom_solver.py:731-739) evaluated here by pandas."""
    import sys
    import types

    import numpy as np  # noqa: F811
    import pandas as pd

    reference_loader.load()
    for name, sub in (("tardis.plasma.properties", "plasma/properties"), ("tardis.opacities.macro_atom", "opacities/macro_atom")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [reference_loader.REF + "/tardis/" + sub]
            sys.modules[name] = m
    from tardis.opacities.macro_atom.macroatom_line_transitions import (
        probability_emission_down,
        probability_internal_down,
        probability_internal_up,
    )
    from tardis.opacities.tau_sobolev import calculate_sobolev_line_opacity, numba_calculate_beta_sobolev

    from . import opacity_oracle

    L, S = len(atomic.nu), plasma.level_number_density.shape[1]
    stim = opacity_oracle.stimulated_emission_factor(atomic, plasma.level_number_density, nlte)
    # one species (Z = 14, ion 1): lines indexed (Z, ion, lower, upper), levels (Z, ion, level) as in atom_data
    lines = pd.DataFrame({"wavelength_cm": atomic.wavelength_cm, "f_lu": atomic.f_lu},
                         index=pd.MultiIndex.from_arrays([np.full(L, 14), np.full(L, 1), atomic.lower_level, atomic.upper_level],
                                                         names=["atomic_number", "ion_number", "level_number_lower", "level_number_upper"]))
    lnd = pd.DataFrame(plasma.level_number_density,
                       index=pd.MultiIndex.from_arrays([np.full(atomic.n_levels, 14), np.full(atomic.n_levels, 1), np.arange(atomic.n_levels)],
                                                       names=["atomic_number", "ion_number", "level_number"]))
    tau = calculate_sobolev_line_opacity(lines, lnd, reference_loader._Q(float(plasma.time_explosion)), stim).to_numpy()
    beta = numba_calculate_beta_sobolev(tau.ravel().copy(), np.empty(tau.size)).reshape(tau.shape)
    nu, f_ul, f_lu = atomic.nu.reshape(-1, 1), atomic.f_ul.reshape(-1, 1), atomic.f_lu.reshape(-1, 1)
    e_lo, e_up = atomic.energy[atomic.lower_level].reshape(-1, 1), atomic.energy[atomic.upper_level].reshape(-1, 1)
    p_em = np.asarray(probability_emission_down(beta, nu, f_ul, e_up, e_lo))
    p_dn = np.asarray(probability_internal_down(beta, nu, f_ul, e_lo))
    p_up = np.asarray(probability_internal_up(beta, nu, f_lu, stim, plasma.j_blues, e_lo))
    rows = atomic.transition_line_idx
    raw = np.where((atomic.transition_type == -1)[:, None], p_em[rows], np.where((atomic.transition_type == 0)[:, None], p_dn[rows], p_up[rows]))
    df = pd.DataFrame(raw)
    df["source"] = atomic.source_block
    norm = df.div(df.groupby("source").transform("sum"))
    norm.replace(np.nan, 0.0, inplace=True)
    norm = norm.drop(columns=["source"]).to_numpy()
    return dict(stimulated_emission_factor=stim, tau_sobolev=tau, beta_sobolev=beta, raw_probabilities=raw, transition_probabilities=norm)


def run_reference_source_function(atomic, tau_sobolev, transition_probabilities, j_blue_estimator, e_dot_lu_estimator, time_explosion,
                                  time_of_simulation, volume, line_interaction_type="macroatom"):
    """att_S_ul, Jred_lu, Jblue_lu, e_dot_u from the UNMODIFIED `SourceFunctionSolver.solve`
    (spectrum/formal_integral/source_function.py:27-143): the pandas frames it reads (atom_data.lines, MacroAtomState's
    transition_metadata / references_index) are built here from `tardis_b200.synthetic.AtomicData` -- one species (Z = 14, ion 1),
    level index == level number -- and the state objects are plain namespaces with the attributes the method reads."""
    import sys
    import types

    import numpy as np  # noqa: F811
    import pandas as pd

    reference_loader.load()
    for name, sub in (("tardis.spectrum", "spectrum"), ("tardis.spectrum.formal_integral", "spectrum/formal_integral")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [reference_loader.REF + "/tardis/" + sub]
            sys.modules[name] = m
    from tardis.spectrum.formal_integral.source_function import SourceFunctionSolver

    L, S = tau_sobolev.shape
    n = atomic.n_levels
    lines = pd.DataFrame({"line_id": np.arange(L), "wavelength_cm": atomic.wavelength_cm},
                         index=pd.MultiIndex.from_arrays([np.full(L, 14), np.full(L, 1), atomic.lower_level, atomic.upper_level],
                                                         names=["atomic_number", "ion_number", "level_number_lower", "level_number_upper"]))
    line = atomic.transition_line_idx
    up = atomic.transition_type == 1
    src = np.where(up, atomic.lower_level[line], atomic.upper_level[line])
    dst = np.where(up, atomic.upper_level[line], atomic.lower_level[line])
    meta = pd.DataFrame({"transition_type": atomic.transition_type, "transition_line_id": line, "source_level_idx": src,
                         "destination_level_idx": dst, "source": [(14, 1, int(s)) for s in src]})
    refs = pd.Series(np.arange(n), index=pd.MultiIndex.from_arrays([np.full(n, 14), np.full(n, 1), np.arange(n)],
                                                                  names=["atomic_number", "ion_number", "level_number"]))
    ns = types.SimpleNamespace
    sim_state = ns(geometry=ns(v_inner_boundary_idx=0, v_outer_boundary_idx=S), no_of_shells=S, dilution_factor=np.ones(S),
                   time_explosion=float(time_explosion), volume=reference_loader._Q(np.asarray(volume, dtype=np.float64)))
    opacity_state = ns(tau_sobolev=np.asarray(tau_sobolev), transition_probabilities=np.asarray(transition_probabilities))
    transport_state = ns(estimators_line=ns(mean_intensity_blueward=np.asarray(j_blue_estimator),
                                            energy_deposition_line_rate=np.asarray(e_dot_lu_estimator)),
                         packet_collection=ns(time_of_simulation=float(time_of_simulation)))
    res = SourceFunctionSolver(line_interaction_type).solve(sim_state, opacity_state, transport_state, ns(lines=lines),
                                                            ns(references_index=refs, transition_metadata=meta))
    levels = np.asarray([ix[2] for ix in res.e_dot_u.index], dtype=np.int64)
    return dict(att_S_ul=np.asarray(res.att_S_ul, dtype=np.float64), Jred_lu=np.asarray(res.Jred_lu, dtype=np.float64),
                Jblue_lu=np.asarray(res.Jblue_lu, dtype=np.float64), e_dot_u=res.e_dot_u.to_numpy(dtype=np.float64), e_dot_u_levels=levels)


def run_reference_formal_integral(model, tau_sobolev, att_S_ul, Jred_lu, Jblue_lu, electron_densities, inner_temperature, frequencies,
                                  points, interpolate_shells):
    """luminosity_densities [n_frequencies] and intensities_nu_p [n_frequencies, points] from the UNMODIFIED
    `FormalIntegralSolver.interpolate_integrator_quantities` (spectrum/formal_integral/formal_integral_solver.py:305-430) followed
    by the UNMODIFIED `numba_formal_integral` (spectrum/formal_integral/formal_integral_numba.py:377-567), glued together the way
    `FormalIntegralSolver.solve` does (:241-285): linspace of `interpolate_shells` radii (0 -> max(2 S, 80); < 0 -> the model's
    shells), tables flattened in Fortran order, a geometry of the interpolated shells.  The interpolated tables come back too."""
    import sys
    import types

    import numpy as np  # noqa: F811
    import pandas as pd

    R = reference_loader.load()
    for name, sub in (("tardis.spectrum", "spectrum"), ("tardis.spectrum.formal_integral", "spectrum/formal_integral")):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [reference_loader.REF + "/tardis/" + sub]
            sys.modules[name] = m
    if "tardis.spectrum.base" not in sys.modules:  # TARDISSpectrum (astropy) is only the return type of FormalIntegralSolver.solve
        sb = types.ModuleType("tardis.spectrum.base")
        sb.TARDISSpectrum = object
        sys.modules["tardis.spectrum.base"] = sb
    from tardis.spectrum.formal_integral.formal_integral_numba import NumbaFormalIntegrator
    from tardis.spectrum.formal_integral.formal_integral_solver import FormalIntegralSolver

    ns = types.SimpleNamespace
    L, S = tau_sobolev.shape
    if interpolate_shells == 0:
        interpolate_shells = max(2 * S, 80)
    r_in, r_out = np.asarray(model.r_inner, dtype=np.float64), np.asarray(model.r_outer, dtype=np.float64)
    if interpolate_shells > 0:
        radius = np.linspace(r_in[0], r_out[-1], interpolate_shells)
        r_in_i, r_out_i = radius[:-1], radius[1:]
    else:
        r_in_i, r_out_i = r_in, r_out
    solver = FormalIntegralSolver(points, interpolate_shells, "numba")
    att_i, jred_i, jblue_i, r_in_i, r_out_i, tau_i, ne_i = solver.interpolate_integrator_quantities(
        r_in, r_out, r_in_i, r_out_i, ns(att_S_ul=np.asarray(att_S_ul), Jred_lu=np.asarray(Jred_lu), Jblue_lu=np.asarray(Jblue_lu)),
        ns(geometry=ns(v_inner_boundary_idx=0, v_outer_boundary_idx=S)), ns(tau_sobolev=pd.DataFrame(np.asarray(tau_sobolev))),
        pd.Series(np.asarray(electron_densities, dtype=np.float64)))
    t_exp = float(model.time_explosion)
    geometry = R.NumbaHomologousRadial1DGeometry(np.ascontiguousarray(r_in_i), np.ascontiguousarray(r_out_i),
                                                 np.ascontiguousarray(r_in_i / t_exp), np.ascontiguousarray(r_out_i / t_exp), t_exp)
    # numba_formal_integral reads only plasma.line_list_nu; a jitclass instance is what the solver passes
    nu_lines = np.ascontiguousarray(model.line_list_nu, dtype=np.float64)
    z1, z2, zi = np.zeros(0), np.zeros((0, 0)), np.zeros(0, dtype=np.int64)
    one = np.zeros((1, 1))
    plasma = R.OpacityStateNumba(np.ascontiguousarray(ne_i, dtype=np.float64), np.zeros(len(ne_i)), nu_lines,
                                 np.ascontiguousarray(tau_i, dtype=np.float64), one, zi, zi, zi, zi, zi,
                                 z1, z2, z1, z1, zi, z2, z1, z1, z1, z2, zi, np.int64(-1))
    integrator = NumbaFormalIntegrator(geometry, t_exp, plasma, points)
    lum, inup = integrator.formal_integral(float(inner_temperature), np.ascontiguousarray(frequencies, dtype=np.float64),
                                           np.ascontiguousarray(att_i.flatten(order="F")), np.ascontiguousarray(jred_i.flatten(order="F")),
                                           np.ascontiguousarray(jblue_i.flatten(order="F")), np.ascontiguousarray(tau_i, dtype=np.float64),
                                           np.ascontiguousarray(ne_i, dtype=np.float64), points)
    return dict(luminosity_densities=np.asarray(lum, dtype=np.float64), intensities_nu_p=np.asarray(inup, dtype=np.float64),
                att_S_ul_interpolated=np.asarray(att_i), Jred_lu_interpolated=np.asarray(jred_i), Jblue_lu_interpolated=np.asarray(jblue_i),
                tau_sobolevs_interpolated=np.asarray(tau_i), electron_densities_interpolated=np.asarray(ne_i),
                r_inner_interpolated=np.asarray(r_in_i), r_outer_interpolated=np.asarray(r_out_i))
