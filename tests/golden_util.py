"""Helpers shared by the golden-vector tests (oracle side and GPU side)."""
import importlib.util
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

_spec = importlib.util.spec_from_file_location("make_golden", os.path.join(GOLDEN_DIR, "make_golden.py"))
make_golden = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(make_golden)

CASES = list(make_golden.CASES)
ORACLE_CASES = CASES + list(make_golden.ORACLE_ONLY_CASES)  # + shapes pinned on the CPU only (their GPU side: bench.py's legs)

# floats: the reference itself is fastmath Numba (not IEEE-reproducible); its own
# regression tolerance for this path is rtol 1e-13 / 1e-12.
RTOL_PACKET = 1e-11
RTOL_EST = 1e-11


def load_case(name):
    """-> (model, packets, run_kwargs, sigma_thomson, golden dict). Asserts the
    regenerated inputs hash to the digest stored when the reference ran."""
    model, packets, rk, sig = make_golden.build_inputs(name)
    g = dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))
    assert make_golden.input_digest(model, packets) == str(g["digest"]), "synthetic inputs drifted from golden"
    return model, packets, rk, sig, g


def oracle_kwargs(rk, sig):
    kw = dict(rk)
    if sig is not None:
        kw["sigma_thomson"] = sig
    return kw


def assert_close(a, b, rtol, name, atol=0.0):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, name
    nan_a, nan_b = np.isnan(a), np.isnan(b)
    assert np.array_equal(nan_a, nan_b), f"{name}: NaN pattern differs"
    if name.endswith("_mu"):
        # direction cosines live in [-1, 1]; mu = (mu r + d)/sqrt(r^2 + d^2 + 2 r d mu) is ill-conditioned for nearly
        # radial inward flights, where the reference (fastmath) and IEEE arithmetic legitimately differ at ~1e-11
        atol = max(atol, 1e-9)
    scale = np.maximum(np.abs(b), np.abs(a))
    err = np.abs(a - b)
    ok = (err <= rtol * scale + atol) | nan_a
    assert ok.all(), f"{name}: max rel err {np.nanmax(err / np.where(scale > 0, scale, 1)):.3e} > {rtol}"


def compare_to_golden(res, g, n_tracked, check_events=True, est_rtol=RTOL_EST, packet_rtol=RTOL_PACKET):
    """`res`: dict from run_oracle or the GPU engine. Integer work exact, floats to tolerance."""
    assert_close(res["output_nus"], g["output_nus"], packet_rtol, "output_nus")
    assert_close(res["output_energies"], g["output_energies"], packet_rtol, "output_energies")
    assert np.array_equal(np.sign(res["output_energies"]), np.sign(g["output_energies"]))
    for k in ("j", "nu_bar", "vhist"):
        assert_close(res[k], g[k], est_rtol, k)
    for k in ("j_blue", "edotlu"):
        if k in g:
            assert_close(res[k], g[k], est_rtol, k)
            # exact zero pattern of the line estimators (lines never passed stay exactly 0)
            assert np.array_equal(res[k] == 0, g[k] == 0)
        else:  # bench-shape goldens: the table is pinned by its checksums (make_golden.compress_line_table)
            c = make_golden.compress_line_table(res[k])
            assert np.array_equal(c["nnz_per_shell"], g[f"{k}__nnz_per_shell"]), f"{k}: non-zero cells per shell differ"
            assert_close(c["bucket_sums"], g[f"{k}__bucket_sums"], est_rtol, f"{k} bucket sums")
            flat = np.asarray(res[k]).ravel()
            assert_close(flat[g[f"{k}__sample_idx"]], g[f"{k}__sample_val"], est_rtol, f"{k} sampled cells")
    if "photo_ion_estimator" in g:  # IIP / continuum mode: EstimatorsContinuum
        for k in ("photo_ion_estimator", "stim_recomb_estimator", "bf_heating_estimator", "stim_recomb_cooling_estimator",
                  "ff_heating_estimator"):
            assert_close(res[k], g[k], est_rtol, k)
        assert np.array_equal(res["photo_ion_estimator_statistics"], g["photo_ion_estimator_statistics"])
    if "last_interaction_type" in res:
        for k in ("last_interaction_type", "last_event_id", "last_shell_id", "last_line_absorb_id", "last_line_emit_id"):
            assert np.array_equal(res[k], g[k]), k
        for k in ("last_radius", "last_before_nu", "last_before_mu", "last_before_energy",
                  "last_after_nu", "last_after_mu", "last_after_energy"):
            assert_close(res[k], g[k], packet_rtol, k)
    if check_events and "events" in res:
        n = min(n_tracked, len(res["events"]))
        assert np.array_equal(res["event_counts"][:n], g["event_counts"][:n]), "event counts differ"
        flat = np.concatenate(res["events"][:n])
        sel = g["ev_packet_id"] < n
        for k in ("packet_id", "interaction_type", "status", "before_shell_id", "after_shell_id",
                  "line_absorb_id", "line_emit_id"):
            assert np.array_equal(flat[k], g["ev_" + k][sel]), f"trajectory field {k} differs"
        for k in ("radius", "before_nu", "before_mu", "before_energy", "after_nu", "after_mu", "after_energy"):
            assert_close(flat[k], g["ev_" + k][sel], packet_rtol, "ev_" + k)
