"""ctypes front-end of the CPU parity oracle (oracle/tardis_oracle.c).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs.  The product package
(tardis_b200/) never imports this.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libtardis_oracle.so")

LINE_INTERACTION = {"scatter": 0, "downbranch": 1, "macroatom": 2}

_pd = C.POINTER(C.c_double)
_pi = C.POINTER(C.c_int64)


class _Model(C.Structure):
    _fields_ = [
        ("n_shells", C.c_int64), ("n_lines", C.c_int64),
        ("r_inner", _pd), ("r_outer", _pd), ("time_explosion", C.c_double),
        ("electron_density", _pd), ("line_list_nu", _pd), ("tau_sobolev", _pd),
        ("n_transitions", C.c_int64), ("n_blocks", C.c_int64),
        ("transition_probabilities", _pd), ("line2macro_level_upper", _pi),
        ("macro_block_edge_index", _pi), ("transition_type", _pi),
        ("destination_level_id", _pi), ("transition_line_id", _pi),
        ("t_electrons", _pd), ("n_continua", C.c_int64), ("n_phot", C.c_int64),
        ("bf_threshold_list_nu", _pd), ("photo_ion_nu_threshold_mins", _pd), ("photo_ion_nu_threshold_maxs", _pd),
        ("photo_ion_block_references", _pi), ("chi_bf", _pd), ("x_sect", _pd), ("phot_nus", _pd),
        ("ff_opacity_factor", _pd), ("emissivities", _pd), ("photo_ion_activation_idx", _pi),
        ("n_activation", C.c_int64), ("k_packet_idx", C.c_int64), ("n_markov", C.c_int64),
        ("absorbing_markov_probabilities", _pd),
    ]


class _Config(C.Structure):
    _fields_ = [
        ("enable_full_relativity", C.c_int), ("line_interaction_type", C.c_int),
        ("disable_line_scattering", C.c_int), ("sigma_thomson", C.c_double),
        ("number_of_vpackets", C.c_int64), ("survival_probability", C.c_double),
        ("vpacket_tau_russian", C.c_double), ("vpacket_spawn_start_frequency", C.c_double),
        ("vpacket_spawn_end_frequency", C.c_double), ("spectrum_frequency_grid", _pd),
        ("n_grid", C.c_int64), ("continuum_processes_enabled", C.c_int),
    ]


class _Packets(C.Structure):
    _fields_ = [
        ("n_packets", C.c_int64), ("initial_radii", _pd), ("initial_nus", _pd),
        ("initial_mus", _pd), ("initial_energies", _pd), ("packet_seeds", _pi),
    ]


class _Counters(C.Structure):
    _fields_ = [(n, C.c_int64) for n in (
        "n_line_steps", "n_boundary_events", "n_line_events", "n_escat_events", "n_rng_draws",
        "n_macro_jumps", "n_macro_scanned", "n_vpackets", "n_vpacket_line_steps", "n_continuum_events",
        "n_bf_estimator_updates")]


EVENT_DTYPE = np.dtype([
    ("packet_id", "i8"), ("interaction_type", "i8"), ("status", "i8"), ("before_shell_id", "i8"),
    ("after_shell_id", "i8"), ("line_absorb_id", "i8"), ("line_emit_id", "i8"),
    ("radius", "f8"), ("before_nu", "f8"), ("before_mu", "f8"), ("before_energy", "f8"),
    ("after_nu", "f8"), ("after_mu", "f8"), ("after_energy", "f8"),
])


class _Outputs(C.Structure):
    _fields_ = [
        ("output_nus", _pd), ("output_energies", _pd), ("j", _pd), ("nu_bar", _pd),
        ("j_blue", _pd), ("edotlu", _pd), ("vhist", _pd),
        ("last_interaction_type", _pi), ("last_event_id", _pi), ("last_shell_id", _pi),
        ("last_line_absorb_id", _pi), ("last_line_emit_id", _pi),
        ("last_radius", _pd), ("last_before_nu", _pd), ("last_before_mu", _pd),
        ("last_before_energy", _pd), ("last_after_nu", _pd), ("last_after_mu", _pd),
        ("last_after_energy", _pd),
        ("events", C.c_void_p), ("event_counts", _pi),
        ("n_tracked_packets", C.c_int64), ("max_events_per_packet", C.c_int64),
        ("vlog_nus", _pd), ("vlog_energies", _pd), ("vlog_initial_mus", _pd), ("vlog_initial_rs", _pd),
        ("vlog_packet_index", _pi), ("vlog_capacity", C.c_int64), ("vlog_count", C.c_int64),
        ("photo_ion_estimator", _pd), ("stim_recomb_estimator", _pd), ("bf_heating_estimator", _pd),
        ("stim_recomb_cooling_estimator", _pd), ("ff_heating_estimator", _pd), ("photo_ion_estimator_statistics", _pi),
        ("counters", _Counters),
    ]


def build(force: bool = False) -> str:
    """Compile the oracle with the recipe in oracle/Makefile (gcc, plain IEEE flags)."""
    deps = [os.path.join(HERE, f) for f in ("tardis_oracle.c", "tardis_oracle.h", "packet_source_oracle.c", "formal_integral_oracle.c", "Makefile")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.run(["make", "-C", HERE, "-B", "libtardis_oracle.so"], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.tardis_oracle_run.restype = C.c_int
        _lib.tardis_oracle_run.argtypes = [C.POINTER(_Model), C.POINTER(_Config), C.POINTER(_Packets),
                                           C.POINTER(_Outputs), C.c_int]
        _lib.tardis_oracle_rng_double.restype = C.c_double
        _lib.tardis_oracle_rng_double.argtypes = [C.c_uint32, C.c_int64]
        _lib.tardis_oracle_distance_boundary.restype = C.c_double
        _lib.tardis_oracle_distance_boundary.argtypes = [C.c_double] * 4 + [_pi]
        _lib.tardis_oracle_distance_line.restype = C.c_double
        _lib.tardis_oracle_distance_line.argtypes = [C.c_double] * 4 + [C.c_int, C.c_double, C.c_double, C.c_int,
                                                                         C.POINTER(C.c_int)]
        _lib.tardis_oracle_doppler_factor.restype = C.c_double
        _lib.tardis_oracle_doppler_factor.argtypes = [C.c_double, C.c_double, C.c_int]
        _lib.tardis_oracle_inverse_doppler_factor.restype = C.c_double
        _lib.tardis_oracle_inverse_doppler_factor.argtypes = [C.c_double, C.c_double, C.c_int]
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_pd)


def _i(a):
    a = np.ascontiguousarray(a, dtype=np.int64)
    return a, a.ctypes.data_as(_pi)


def rng_double(seed: int, skip: int = 0) -> float:
    return lib().tardis_oracle_rng_double(seed & 0xFFFFFFFF, skip)


def distance_boundary(r, mu, r_inner, r_outer):
    ds = C.c_int64(0)
    d = lib().tardis_oracle_distance_boundary(r, mu, r_inner, r_outer, C.byref(ds))
    return d, ds.value


def distance_line(r, mu, nu, comov_nu, is_last_line, nu_line, time_explosion, full_rel=False):
    err = C.c_int(0)
    d = lib().tardis_oracle_distance_line(r, mu, nu, comov_nu, int(is_last_line), nu_line, time_explosion,
                                          int(full_rel), C.byref(err))
    return d, err.value


def doppler_factor(velocity, mu, full_rel=False):
    return lib().tardis_oracle_doppler_factor(velocity, mu, int(full_rel))


def inverse_doppler_factor(velocity, mu, full_rel=False):
    return lib().tardis_oracle_inverse_doppler_factor(velocity, mu, int(full_rel))


class OracleError(RuntimeError):
    pass


def run_oracle(model, packets, *, number_of_vpackets=0, enable_full_relativity=False,
               disable_line_scattering=False, survival_probability=0.0, vpacket_tau_russian=10.0,
               spawn_start=0.0, spawn_end=1e200, sigma_thomson=6.652458734e-25,
               n_tracked_packets=0, max_events_per_packet=512, vlog_capacity=0,
               track_last_interaction=True, nthreads=1, private_tables_max_threads=None):
    """One MC iteration on the CPU oracle.  Returns a dict with the same keys as
    oracle.reference_runner.run_reference (plus `counters`, `events`)."""
    L = lib()
    keep = []
    m = _Model()
    m.n_shells = model.n_shells
    m.n_lines = model.n_lines
    for name in ("r_inner", "r_outer", "electron_density", "line_list_nu", "tau_sobolev"):
        a, p = _d(getattr(model, name)); keep.append(a); setattr(m, name, p)
    m.time_explosion = float(model.time_explosion)
    mac = model.macro
    a, p = _d(mac.transition_probabilities); keep.append(a); m.transition_probabilities = p
    m.n_transitions = mac.transition_probabilities.shape[0]
    m.n_blocks = len(mac.macro_block_edge_index) - 1
    for name in ("line2macro_level_upper", "macro_block_edge_index", "transition_type",
                 "destination_level_id", "transition_line_id"):
        a, p = _i(getattr(mac, name)); keep.append(a); setattr(m, name, p)

    cont = getattr(model, "continuum", None)
    if cont is not None:
        # IIP mode: the macro-atom tables are the continuum-aware ones (normalized deactivation probabilities)
        a, p = _d(model.t_electrons); keep.append(a); m.t_electrons = p
        m.n_continua = len(cont.bf_threshold_list_nu)
        m.n_phot = len(cont.phot_nus)
        for name in ("bf_threshold_list_nu", "photo_ion_nu_threshold_mins", "photo_ion_nu_threshold_maxs", "chi_bf",
                     "x_sect", "phot_nus", "ff_opacity_factor", "emissivities", "absorbing_markov_probabilities"):
            a, p = _d(getattr(cont, name)); keep.append(a); setattr(m, name, p)
        for name in ("photo_ion_block_references", "photo_ion_activation_idx"):
            a, p = _i(getattr(cont, name)); keep.append(a); setattr(m, name, p)
        m.n_activation = len(cont.photo_ion_activation_idx)
        m.k_packet_idx = int(cont.k_packet_idx)
        m.n_markov = cont.absorbing_markov_probabilities.shape[1]

    c = _Config()
    c.continuum_processes_enabled = int(cont is not None)
    c.enable_full_relativity = int(enable_full_relativity)
    c.line_interaction_type = LINE_INTERACTION[model.line_interaction_type]
    c.disable_line_scattering = int(disable_line_scattering)
    c.sigma_thomson = sigma_thomson
    c.number_of_vpackets = number_of_vpackets
    c.survival_probability = survival_probability
    c.vpacket_tau_russian = vpacket_tau_russian
    c.vpacket_spawn_start_frequency = spawn_start
    c.vpacket_spawn_end_frequency = spawn_end
    a, p = _d(model.spectrum_frequency_grid); keep.append(a); c.spectrum_frequency_grid = p
    c.n_grid = len(model.spectrum_frequency_grid)

    pk = _Packets()
    n = len(packets)
    pk.n_packets = n
    for name in ("initial_radii", "initial_nus", "initial_mus", "initial_energies"):
        a, p = _d(getattr(packets, name)); keep.append(a); setattr(pk, name, p)
    a, p = _i(packets.packet_seeds); keep.append(a); pk.packet_seeds = p

    S, Ln = model.n_shells, model.n_lines
    res = dict(
        output_nus=np.empty(n), output_energies=np.empty(n), j=np.zeros(S), nu_bar=np.zeros(S),
        j_blue=np.zeros((Ln, S)), edotlu=np.zeros((Ln, S)), vhist=np.zeros(c.n_grid),
    )
    o = _Outputs()
    for k, v in res.items():
        setattr(o, k, v.ctypes.data_as(_pd))
    if track_last_interaction:
        for k in ("last_interaction_type", "last_event_id", "last_shell_id", "last_line_absorb_id", "last_line_emit_id"):
            res[k] = np.empty(n, dtype=np.int64); setattr(o, k, res[k].ctypes.data_as(_pi))
        for k in ("last_radius", "last_before_nu", "last_before_mu", "last_before_energy",
                  "last_after_nu", "last_after_mu", "last_after_energy"):
            res[k] = np.empty(n); setattr(o, k, res[k].ctypes.data_as(_pd))
    if cont is not None:
        nc = len(cont.bf_threshold_list_nu)
        for k in ("photo_ion_estimator", "stim_recomb_estimator", "bf_heating_estimator", "stim_recomb_cooling_estimator"):
            res[k] = np.zeros((nc, S)); setattr(o, k, res[k].ctypes.data_as(_pd))
        res["ff_heating_estimator"] = np.zeros(S); o.ff_heating_estimator = res["ff_heating_estimator"].ctypes.data_as(_pd)
        res["photo_ion_estimator_statistics"] = np.zeros((nc, S), dtype=np.int64)
        o.photo_ion_estimator_statistics = res["photo_ion_estimator_statistics"].ctypes.data_as(_pi)
    n_tracked_packets = min(n_tracked_packets, n)
    if n_tracked_packets > 0:
        ev = np.zeros(n_tracked_packets * max_events_per_packet, dtype=EVENT_DTYPE)
        evc = np.zeros(n_tracked_packets, dtype=np.int64)
        o.events = ev.ctypes.data
        o.event_counts = evc.ctypes.data_as(_pi)
        o.n_tracked_packets = n_tracked_packets
        o.max_events_per_packet = max_events_per_packet
    if vlog_capacity > 0:
        for k in ("vlog_nus", "vlog_energies", "vlog_initial_mus", "vlog_initial_rs"):
            res[k] = np.zeros(vlog_capacity); setattr(o, k, res[k].ctypes.data_as(_pd))
        res["vlog_packet_index"] = np.zeros(vlog_capacity, dtype=np.int64)
        o.vlog_packet_index = res["vlog_packet_index"].ctypes.data_as(_pi)
        o.vlog_capacity = vlog_capacity
    if private_tables_max_threads is not None:
        os.environ["TARDIS_ORACLE_PRIVATE_MAX"] = str(int(private_tables_max_threads))
    else:
        os.environ.pop("TARDIS_ORACLE_PRIVATE_MAX", None)
    err = L.tardis_oracle_run(C.byref(m), C.byref(c), C.byref(pk), C.byref(o), int(nthreads))
    if err:
        raise OracleError({1: "nu difference is less than 0.0", 2: "MacroAtomError",
                           3: "vpacket did not terminate", 4: "continuum tables inconsistent"}.get(err, str(err)))
    res["counters"] = {k: getattr(o.counters, k) for k, _ in _Counters._fields_}
    if n_tracked_packets > 0:
        ev = ev.reshape(n_tracked_packets, max_events_per_packet)
        res["events"] = [ev[i, : evc[i]] for i in range(n_tracked_packets)]
        res["event_counts"] = evc
    if vlog_capacity > 0:
        res["vlog_count"] = o.vlog_count
    return res


def create_packets(n_packets: int, seed: int, radius: float, temperature: float, l_samples: int = 1000,
                   max_seed_val: int = 2**32 - 1, beta: float | None = None):
    """Sequential restatement of BlackBodySimpleSource.create_packets for np.random.default_rng(seed)
    (oracle/packet_source_oracle.c).  Returns a dict with the PacketCollection arrays."""
    L = lib()
    L.tardis_oracle_create_packets_beta.restype = C.c_int
    L.tardis_oracle_create_packets_beta.argtypes = [C.c_uint64, C.c_int64, C.c_uint32, C.c_double, C.c_double, C.c_double, C.c_double,
                                                    _pd, C.c_int64, C.c_double, _pd, _pd, _pd, _pd, _pd, _pi, C.c_double]
    n = int(n_packets)
    l_array = np.cumsum(np.arange(1, l_samples, dtype=np.float64) ** -4)
    out = {k: np.empty(n, dtype=np.float64) for k in ("initial_radii", "initial_nus", "initial_mus", "initial_energies")}
    out["packet_seeds"] = np.empty(n, dtype=np.int64)
    scratch = np.empty(5 * n, dtype=np.float64)
    err = L.tardis_oracle_create_packets_beta(int(seed), n, int(max_seed_val), float(radius), float(temperature), 1.3806488e-16, 6.62606957e-27,
                                         l_array.ctypes.data_as(_pd), len(l_array), float(np.pi**4 / 90.0), scratch.ctypes.data_as(_pd),
                                         out["initial_radii"].ctypes.data_as(_pd), out["initial_nus"].ctypes.data_as(_pd),
                                         out["initial_mus"].ctypes.data_as(_pd), out["initial_energies"].ctypes.data_as(_pd),
                                         out["packet_seeds"].ctypes.data_as(_pi), -1.0 if beta is None else float(beta))
    if err:
        raise ValueError(f"tardis_oracle_create_packets: error {err}")
    return out
