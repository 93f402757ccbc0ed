"""Timing (and a parity spot check) of tb200_formal_integral at the bench shape: 5e5 lines, 20 -> 79 shells, the spectrum grid's
10 000 frequencies x 1000 impact parameters.  Tables given from the host (random, physically ordered), so the probe needs no
transport run.  python scripts/formal_integral_probe.py [--frequencies N] [--points P] [--check K] [--repeat R]"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tardis_b200 import synthetic as syn  # noqa: E402
from tardis_b200.engine import Engine  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--lines", type=int, default=500_000)
    ap.add_argument("--shells", type=int, default=20)
    ap.add_argument("--frequencies", type=int, default=10_000)
    ap.add_argument("--points", type=int, default=1000)
    ap.add_argument("--interpolate-shells", type=int, default=0)
    ap.add_argument("--check", type=int, default=2, help="frequencies compared with the CPU oracle (0: none)")
    ap.add_argument("--repeat", type=int, default=3)
    args = ap.parse_args()
    L, S = args.lines, args.shells
    model = syn.make_model(S, L, "downbranch", mu_tau=-3.0)
    m = model.macro
    rng = np.random.default_rng(5)
    att = rng.random((L, S)) * 1e-6
    jblue = rng.random((L, S)) * 1e-5
    jred = jblue * np.exp(-np.asarray(model.tau_sobolev)) + att
    eng = Engine(0)
    eng.set_model(r_inner=model.r_inner, r_outer=model.r_outer, time_explosion=model.time_explosion, electron_density=model.electron_density,
                  line_list_nu=model.line_list_nu, tau_sobolev=model.tau_sobolev, line_interaction_type="downbranch",
                  transition_probabilities=m.transition_probabilities, line2macro_level_upper=m.line2macro_level_upper,
                  macro_block_edge_index=m.macro_block_edge_index, transition_type=m.transition_type,
                  destination_level_id=m.destination_level_id, transition_line_id=m.transition_line_id,
                  spectrum_frequency_grid=model.spectrum_frequency_grid)
    grid = np.asarray(model.spectrum_frequency_grid, dtype=np.float64)
    freq = np.linspace(grid[0], grid[-2], args.frequencies)
    runs = []
    res = None
    for _ in range(args.repeat):
        t0 = time.perf_counter()
        res = eng.formal_integral(inner_temperature=1e4, frequencies=freq, points=args.points, interpolate_shells=args.interpolate_shells,
                                  tables=(att, jred, jblue))
        runs.append({"wall_ms": (time.perf_counter() - t0) * 1e3, "interpolation_ms": res["interpolation_ms"], "integral_ms": res["integral_ms"]})
    best = min(r["integral_ms"] for r in runs)
    out = {"n_lines": L, "n_shells": S, "n_frequencies": args.frequencies, "n_impact_parameters": args.points,
           "interpolate_shells": args.interpolate_shells, "runs": runs, "frequencies_per_s": args.frequencies / (best * 1e-3),
           "rays_per_s": args.frequencies * (args.points - 1) / (best * 1e-3)}
    if args.check > 0:
        from oracle import formal_integral_oracle as fio  # the checker, not the thing measured

        sample = np.linspace(0, len(freq) - 1, args.check + 2).astype(int)[1:-1]
        t0 = time.perf_counter()
        want = fio.solve(model.r_inner, model.r_outer, float(model.time_explosion), model.line_list_nu, 1e4, freq[sample], att, jred, jblue,
                         model.tau_sobolev, model.electron_density, args.points, args.interpolate_shells)
        out["oracle_s"] = time.perf_counter() - t0
        got = res["luminosity_densities"][sample]
        out["max_rel_err_L_nu_vs_oracle"] = float(np.max(np.abs(got - want["luminosity_densities"]) / np.abs(want["luminosity_densities"])))
        out["L_nu_sample"] = [float(x) for x in got]
    print(json.dumps(out))
    eng.close()


main()
