"""Run one golden case on the GPU engine and report (diagnostics; used while bringing kernels up)."""
import sys
import time

import numpy as np

sys.path.insert(0, "tests")
sys.path.insert(0, ".")
from golden_util import compare_to_golden, load_case, make_golden, oracle_kwargs  # noqa: E402
from test_gpu_parity import engine_config  # noqa: E402


def main():
    name, algo = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0
    t0 = time.time()
    from tardis_b200.engine import Engine

    model, packets, rk, sig, g = load_case(name)
    eng = Engine(0)
    eng.set_option("algorithm", algo)
    eng.set_model_from(model, **engine_config(rk, sig))
    print(name, "algo", algo, "model set", round(time.time() - t0, 2), flush=True)
    try:
        res = eng.run_packets(packets, track_last_interaction=True, n_tracked_packets=make_golden.N_TRACKED, max_events_per_packet=4096)
    except Exception as e:  # noqa: BLE001
        print("ENGINE ERROR", type(e).__name__, str(e)[:300], flush=True)
        return
    print("ran; kernel ms", eng.last_kernel_ms(), res["counters"], flush=True)
    from oracle import cpu_oracle

    ref = cpu_oracle.run_oracle(model, packets, **oracle_kwargs(rk, sig))
    print("oracle counters", ref["counters"], flush=True)
    try:
        compare_to_golden(res, g, make_golden.N_TRACKED)
        print("PARITY OK", name, flush=True)
    except AssertionError as e:
        print("PARITY FAIL", name, str(e)[:500], flush=True)
        bad = np.nonzero(~np.isclose(res["output_nus"], g["output_nus"], rtol=1e-9, atol=0))[0]
        print("n bad packets", len(bad), bad[:10], flush=True)


if __name__ == "__main__":
    main()
