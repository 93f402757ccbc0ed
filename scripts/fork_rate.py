"""Fork rate of the jump algorithm against the oracle at the bench shape (VERDICT r01, weak #1 / SURVEY.md §7).

The jump kernels decide where a trace ends with arithmetic that is equivalent to, but not literally, the reference's
(double-double prefix differences instead of a running fp64 sum, `* inv_chi` instead of `/ chi`, frequency windows instead of
distances in full relativity ...): a decision can fall the other way when two quantities tie to the last bit.  This script
counts how often: N packets of the bench model (5e5 lines, 20 shells, macroatom) through the engine and through the oracle
(oracle/tardis_oracle.c, pinned against the reference at this very shape: tests/golden/bench_macroatom.npz), then every
per-packet integer (fate, last interaction type / shell / absorbed line / emitted line, interaction count) and the eleven
work counters are compared.  A packet with ANY differing integer is a fork.

    python scripts/fork_rate.py [--packets 10000000] [--out profiles/r02_fork_rate.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from oracle import cpu_oracle
from tardis_b200 import synthetic as syn
from tardis_b200.engine import Engine


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--packets", type=int, default=10_000_000)
    ap.add_argument("--continuum", action="store_true")
    ap.add_argument("--shells", type=int, default=20)
    ap.add_argument("--vpackets", type=int, default=0)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    model = syn.make_model(args.shells, 500_000, "macroatom", mu_tau=-7.5)
    if args.continuum:
        syn.add_continuum(model)
    parts = [syn.make_packets(min(5_000_000, args.packets - lo), model.r_inner[0], base_seed=syn.BASE_SEED + 4242, iteration=i)
             for i, lo in enumerate(range(0, args.packets, 5_000_000))]
    packets = syn.Packets(*(np.concatenate([getattr(q, f) for q in parts]) for f in
                            ("initial_radii", "initial_nus", "initial_mus", "initial_energies", "packet_seeds")), parts[0].radiation_field_luminosity)
    packets.initial_energies[:] = 1.0 / len(packets)
    n = len(packets)
    eng = Engine(0)
    eng.set_model_from(model, number_of_vpackets=args.vpackets)
    t0 = time.perf_counter()
    g = eng.run_packets(packets, track_last_interaction=True)
    t_gpu = time.perf_counter() - t0
    threads = os.cpu_count() or 1
    t0 = time.perf_counter()
    r = cpu_oracle.run_oracle(model, packets, number_of_vpackets=args.vpackets, nthreads=threads, track_last_interaction=True,
                              private_tables_max_threads=0)
    t_cpu = time.perf_counter() - t0
    int_cols = ("last_interaction_type", "last_event_id", "last_shell_id", "last_line_absorb_id", "last_line_emit_id")
    differ = np.sign(g["output_energies"]) != np.sign(r["output_energies"])
    per_col = {"fate": int(differ.sum())}
    for k in int_cols:
        d = g[k] != r[k]
        per_col[k] = int(d.sum())
        differ |= d
    forks = int(differ.sum())
    same = ~differ

    def relerr(a, b):
        a, b = a[same], b[same]
        s = np.maximum(np.abs(a), np.abs(b))
        return float(np.max(np.abs(a - b) / np.where(s > 0, s, 1.0)))

    counters_equal = {k: (int(g["counters"][k]), int(v)) for k, v in r["counters"].items() if g["counters"][k] != v}
    est = {}
    for k in ("j", "nu_bar", "j_blue", "edotlu"):
        nz = r[k] != 0
        est[k] = {"max_rel_err": float(np.max(np.abs(g[k][nz] - r[k][nz]) / np.abs(r[k][nz]))), "zero_pattern_equal": bool(np.array_equal(g[k] == 0, r[k] == 0))}
    out = {"what": "jump algorithm (engine default) vs oracle, bench model: 5e5 lines, %d shells, macroatom%s%s" %
                   (args.shells, ", continuum" if args.continuum else "", f", {args.vpackets} vpackets" if args.vpackets else ""),
           "packets": n, "forked_packets": forks, "fork_rate": forks / n, "differing_integers_by_column": per_col,
           "work_counters_that_differ": counters_equal, "per_packet_float_max_rel_err_of_unforked": {
               "output_nus": relerr(g["output_nus"], r["output_nus"]), "output_energies": relerr(g["output_energies"], r["output_energies"])},
           "estimators": est, "seconds": {"engine_run_incl_copies": t_gpu, "oracle": t_cpu, "oracle_threads": threads}}
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
