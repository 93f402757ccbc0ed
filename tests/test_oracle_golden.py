"""Pin the CPU oracle (oracle/tardis_oracle.c) against golden vectors produced by the
UNMODIFIED reference Numba path (tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest

from golden_util import ORACLE_CASES, compare_to_golden, load_case, make_golden, oracle_kwargs


@pytest.mark.parametrize("name", ORACLE_CASES)
def test_oracle_matches_reference_golden(oracle, name):
    model, packets, rk, sig, g = load_case(name)
    res = oracle.run_oracle(model, packets, n_tracked_packets=make_golden.N_TRACKED, max_events_per_packet=4096,
                            **oracle_kwargs(rk, sig))
    compare_to_golden(res, g, make_golden.N_TRACKED)


def test_oracle_multithreaded_matches_single(oracle):
    model, packets, rk, sig, g = load_case("macroatom_vpackets")
    a = oracle.run_oracle(model, packets, nthreads=1, **oracle_kwargs(rk, sig))
    b = oracle.run_oracle(model, packets, nthreads=4, **oracle_kwargs(rk, sig))
    assert np.array_equal(a["output_nus"], b["output_nus"])
    assert np.array_equal(a["output_energies"], b["output_energies"])
    assert a["counters"] == b["counters"]
    for k in ("j", "nu_bar", "j_blue", "edotlu", "vhist"):
        np.testing.assert_allclose(a[k], b[k], rtol=1e-12, atol=0)
