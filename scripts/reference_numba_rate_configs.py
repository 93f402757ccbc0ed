"""Packets/s of the UNMODIFIED reference Numba loops on the models of BASELINE configs 2, 4 and 5 (config 3: reference_numba_rate.py),
in THIS container: scatter / macroatom + 10 virtual packets (`montecarlo_transport_with_vpackets`, modes/montecarlo_transport.py:239) and
the continuum (IIP) mode (`montecarlo_transport`, modes/iip/montecarlo_transport.py:40).  One warm-up call per thread count (JIT), then
best of 2; the timed region is the reference call itself ("loop only") and, beside it, with the per-packet tracker list it needs.
Each config runs in its own process (the reference freezes CONTINUUM_PROCESSES_ENABLED into the code it compiles first).

    python scripts/reference_numba_rate_configs.py [--out profiles/r02_reference_numba_rate_configs.json]
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {"2": dict(shells=20, mode="scatter", vpackets=0, continuum=False, packets=100_000),
           "4": dict(shells=20, mode="macroatom", vpackets=10, continuum=False, packets=20_000),
           "5": dict(shells=50, mode="macroatom", vpackets=0, continuum=True, packets=20_000)}


def one(key):
    import numba

    from oracle import reference_runner
    from tardis_b200 import synthetic as syn

    c = CONFIGS[key]
    model = syn.make_model(c["shells"], 500_000, c["mode"], mu_tau=-7.5)
    if c["continuum"]:
        syn.add_continuum(model)
    packets = syn.make_packets(c["packets"], model.r_inner[0], base_seed=syn.BASE_SEED + 777)
    out = {}
    for nthreads in (os.cpu_count() or 1, 1):
        best = None
        for rep in range(3):  # rep 0: JIT warm-up
            t = {}
            if c["continuum"]:
                res = reference_runner.run_reference_iip(model, packets, nthreads=nthreads, timings=t)
            else:
                res = reference_runner.run_reference(model, packets, number_of_vpackets=c["vpackets"], nthreads=nthreads, timings=t)
            if rep > 0 and (best is None or t["loop_s"] < best["loop_s"]):
                best = t
        out[str(nthreads)] = {"threads": nthreads, "packets_per_s_loop_only": c["packets"] / best["loop_s"],
                              "packets_per_s_incl_tracker_setup": c["packets"] / (best["loop_s"] + best["tracker_setup_s"]), "seconds_loop_only": best["loop_s"]}
    emitted = float((res["output_energies"] >= 0).mean())
    print(json.dumps({"config": key, **c, "emitted_fraction": emitted, "numba": numba.__version__, "results": out}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--one", default=None)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    if args.one:
        one(args.one)
        return
    rows = {}
    for key in CONFIGS:
        p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", key], capture_output=True, text=True, cwd=ROOT)
        if p.returncode != 0:
            raise SystemExit(p.stderr[-3000:])
        rows[key] = json.loads([ln for ln in p.stdout.splitlines() if ln.startswith("{")][-1])
        print(key, rows[key]["results"], flush=True)
    out = {"what": "unmodified reference Numba loops on the bench models of BASELINE configs 2 / 4 / 5 (5e5 lines, tau~10^N(-7.5,2))",
           "host": {"cpu_count": os.cpu_count(), "where": "build container (no GPU)"}, "configs": rows}
    if args.out:
        json.dump(out, open(args.out, "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
