"""Quick device-side timing of the transport kernel on a synthetic model (not the bench)."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np

from tardis_b200 import synthetic as syn
from tardis_b200.engine import Engine


def alg_bytes(c, n):
    return (48 * c["n_line_steps"] + 16 * c["n_vpacket_line_steps"]
            + 32 * (c["n_boundary_events"] + c["n_line_events"] + c["n_escat_events"])
            + 8 * c["n_macro_scanned"] + 24 * c["n_macro_jumps"] + 56 * n)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--packets", type=int, default=2_000_000)
    ap.add_argument("--lines", type=int, default=500_000)
    ap.add_argument("--shells", type=int, default=20)
    ap.add_argument("--mode", default="scatter")
    ap.add_argument("--mu-tau", type=float, default=-7.5)
    ap.add_argument("--vpackets", type=int, default=0)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--configs", default="2x256")
    ap.add_argument("--algorithm", type=int, default=0)
    ap.add_argument("--continuum", action="store_true", help="IIP mode: add synthetic continuum tables")
    ap.add_argument("--opt", action="append", default=[], help="name=value engine options")
    args = ap.parse_args()
    t0 = time.time()
    model = syn.make_model(args.shells, args.lines, args.mode, mu_tau=args.mu_tau)
    if args.continuum:
        syn.add_continuum(model)
    packets = syn.make_packets(args.packets, model.r_inner[0])
    print(f"# model+packets built in {time.time()-t0:.1f}s", flush=True)
    eng = Engine(0)
    eng.set_option("algorithm", args.algorithm)
    for o in args.opt:
        k, v = o.split("=")
        eng.set_option(k, int(v))
    eng.set_model_from(model, number_of_vpackets=args.vpackets)
    eng.upload_packets(packets.initial_radii, packets.initial_nus, packets.initial_mus, packets.initial_energies,
                       packets.packet_seeds)
    for cfg in args.configs.split(","):
        ctas, threads = (int(x) for x in cfg.split("x"))
        eng.set_option("ctas_per_sm", ctas)
        eng.set_option("threads_per_cta", threads)
        best = None
        for rep in range(args.reps):
            eng.transport(True)
            eng.sync()
            ms = eng.last_kernel_ms()
            best = ms if best is None else min(best, ms)
        c = eng.counters()
        n = args.packets
        ab = alg_bytes(c, n)
        print(json.dumps(dict(config=cfg, ms=round(best, 3), packets_per_s=round(n / best * 1e3), line_steps_per_packet=round(c["n_line_steps"] / n, 1),
                              events_per_packet=round((c["n_boundary_events"] + c["n_line_events"] + c["n_escat_events"]) / n, 2),
                              continuum_events_per_packet=round(c.get("n_continuum_events", 0) / n, 2),
                              alg_GBps=round(ab / best / 1e6, 1), frac_of_6562=round(ab / best / 1e6 / 6562.6, 3), counters=c)), flush=True)


if __name__ == "__main__":
    main()
