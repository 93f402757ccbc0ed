"""Frequencies/s of the UNMODIFIED reference formal integral (`FormalIntegralSolver.interpolate_integrator_quantities` +
`numba_formal_integral`, tardis/spectrum/formal_integral/formal_integral_solver.py:305-430, formal_integral_numba.py:377-567) at
the bench shape -- 5e5 lines, 20 -> 79 shells, 1000 impact parameters -- in THIS container (the reference cannot travel to the
GPU box), and a golden of its output at that shape for tests/test_formal_integral.py (the oracle's pin at full size).

    python scripts/reference_formal_integral_rate.py [--frequencies 16] [--out profiles/r02_reference_formal_integral_rate.json]
                                                     [--golden tests/golden/formal_integral_bench_shape.npz]

Inputs are regenerated from seeds by tests/golden/make_golden.py::formal_integral_bench_shape_inputs (the tests call it too), so the
golden holds only outputs."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np


sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from golden_util import make_golden  # noqa: E402  (the inputs live next to the golden they belong to)

L, S = make_golden.FORMAL_INTEGRAL_BENCH_SHAPE["n_lines"], make_golden.FORMAL_INTEGRAL_BENCH_SHAPE["n_shells"]
POINTS, T_INNER = make_golden.FORMAL_INTEGRAL_BENCH_SHAPE["points"], make_golden.FORMAL_INTEGRAL_BENCH_SHAPE["inner_temperature"]


def bench_shape_inputs(n_frequencies: int):
    i = make_golden.formal_integral_bench_shape_inputs(n_frequencies)
    return i["model"], i["att_S_ul"], i["Jred_lu"], i["Jblue_lu"], i["frequencies"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frequencies", type=int, default=16)
    ap.add_argument("--out", default=None)
    ap.add_argument("--golden", default=None)
    args = ap.parse_args()
    import numba

    from oracle import reference_runner

    model, att, jred, jblue, freq = bench_shape_inputs(args.frequencies)
    # phase timers around the reference's own two calls (wrapped, not modified)
    phase = {"interpolation": [], "integral": []}
    reference_runner.run_reference_formal_integral(model, np.asarray(model.tau_sobolev), att, jred, jblue, model.electron_density, T_INNER, freq[:1],
                                                   POINTS, 0)  # imports + JIT warm-up
    fis = sys.modules["tardis.spectrum.formal_integral.formal_integral_solver"].FormalIntegralSolver
    nfi = sys.modules["tardis.spectrum.formal_integral.formal_integral_numba"].NumbaFormalIntegrator

    def timed(cls, name, key):
        inner = getattr(cls, name)

        def wrapper(*a, **k):
            t0 = time.perf_counter()
            out = inner(*a, **k)
            phase[key].append(time.perf_counter() - t0)
            return out

        setattr(cls, name, wrapper)

    timed(fis, "interpolate_integrator_quantities", "interpolation")
    timed(nfi, "formal_integral", "integral")
    results, ref = {}, None
    for nthreads in (os.cpu_count() or 1, 1):
        numba.set_num_threads(nthreads)
        phase["interpolation"].clear(); phase["integral"].clear()
        for rep in range(3):
            ref = reference_runner.run_reference_formal_integral(model, np.asarray(model.tau_sobolev), att, jred, jblue, model.electron_density,
                                                                 T_INNER, freq, POINTS, 0)
            print(nthreads, rep, "interpolation %.2f s, integral of %d frequencies %.2f s" % (phase["interpolation"][-1], len(freq), phase["integral"][-1]),
                  flush=True)
        results[str(nthreads)] = {"threads": nthreads, "seconds_interpolation": min(phase["interpolation"]), "seconds_integral": min(phase["integral"]),
                                  "frequencies_per_s_integral_only": len(freq) / min(phase["integral"])}
    t_interp = min(r["seconds_interpolation"] for r in results.values())
    out = {"what": "unmodified reference: interpolate_integrator_quantities (scipy interp1d, four [L,79] tables) + numba_formal_integral "
                   "(prange over frequencies), bench shape; best of 3 after a JIT warm-up, the two calls timed separately",
           "n_lines": L, "n_shells": S, "integrator_shells": 79, "n_impact_parameters": POINTS,
           "frequencies": len(freq), "seconds_interpolation": t_interp,
           "seconds_for_the_reference_grid_of_10000_frequencies": {k: t_interp + 10000 / r["frequencies_per_s_integral_only"] for k, r in results.items()},
           "host": {"cpu_count": os.cpu_count(), "numba": numba.__version__, "where": "build container (no GPU)"}, "results": results}
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)
    if args.golden:
        np.savez_compressed(args.golden, frequencies=freq, luminosity_densities=ref["luminosity_densities"], intensities_nu_p=ref["intensities_nu_p"],
                            n_lines=L, n_shells=S, points=POINTS, inner_temperature=T_INNER)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
