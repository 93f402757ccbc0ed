"""Packets/s of the UNMODIFIED reference Numba loop (`montecarlo_transport_with_vpackets`,
tardis/transport/montecarlo/modes/montecarlo_transport.py:239) on the bench model, in THIS container
(the reference cannot travel to the GPU box: no /root/reference there).  BASELINE.md §3 protocol: one warm-up call
(JIT), then best of 3; numba.set_num_threads(1) and (all cores).  "loop only" is the call itself; "incl. setup" adds the
per-packet tracker list the reference builds for every iteration (modes/classic/solver.py:209-211).

    python scripts/reference_numba_rate.py [--packets 100000] [--out profiles/r02_reference_numba_rate.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from oracle import reference_loader, reference_runner
from tardis_b200 import synthetic as syn


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--packets", type=int, default=100_000)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import numba

    model = syn.make_model(20, 500_000, "macroatom", mu_tau=-7.5)
    packets = syn.make_packets(args.packets, model.r_inner[0], base_seed=syn.BASE_SEED + 777)
    R, geometry, opacity, cfg, pc = reference_runner.build_reference_objects(model, packets)
    n = len(packets)
    results = {}
    for nthreads in (1, os.cpu_count() or 1):
        numba.set_num_threads(nthreads)
        best_loop, best_total = None, None
        for rep in range(4):  # rep 0 = JIT warm-up
            t0 = time.perf_counter()
            trackers = R.generate_tracker_last_interaction_list(n)
            t1 = time.perf_counter()
            R.montecarlo_transport_with_vpackets(pc, geometry, model.time_explosion, opacity, cfg, model.spectrum_frequency_grid,
                                                 trackers, 0, False, R.packet_propagation)
            t2 = time.perf_counter()
            if rep > 0:
                best_loop = (t2 - t1) if best_loop is None else min(best_loop, t2 - t1)
                best_total = (t2 - t0) if best_total is None else min(best_total, t2 - t0)
        results[str(nthreads)] = {"threads": nthreads, "packets_per_s_loop_only": n / best_loop, "packets_per_s_incl_tracker_setup": n / best_total,
                                  "seconds_loop_only": best_loop}
        print(nthreads, results[str(nthreads)], flush=True)
    emitted = float((np.asarray(pc.output_energies) >= 0).mean())
    out = {"what": "unmodified reference Numba loop, bench model (5e5 lines, 20 shells, macroatom, tau~10^N(-7.5,2)), no virtual packets",
           "packets": n, "host": {"cpu_count": os.cpu_count(), "numba": numba.__version__, "where": "build container (no GPU)"},
           "emitted_fraction": emitted, "results": results}
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
