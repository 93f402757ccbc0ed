"""Known-answer tests the reference's own test-suite holds inline for this path
(SURVEY.md §8c), applied to the CPU oracle.  No external data needed."""
import numpy as np
import pytest

C = 2.99792458e10


def test_get_random_mu_seed_1963(oracle):
    # tardis/transport/montecarlo/packets/tests/test_packet.py:162-169
    assert 2.0 * oracle.rng_double(1963) - 1.0 == 0.9136407866175174


@pytest.mark.parametrize(
    "r, mu, expected_d, expected_shell",
    [
        (7.5e14, 0.3, 259376919351035.88, 1),
        (7.5e14, -0.3, 709376919351035.9, 1),  # inward but misses the inner boundary
    ],
)
def test_calculate_distance_boundary(oracle, r, mu, expected_d, expected_shell):
    # packets/tests/test_packet.py:65-82 (r_inner=6.912e14, r_outer=8.64e14)
    d, ds = oracle.distance_boundary(r, mu, 6.912e14, 8.64e14)
    np.testing.assert_allclose(d, expected_d, rtol=1e-15)
    assert ds == expected_shell


def test_calculate_distance_boundary_inside_inner(oracle):
    # packets/tests/test_packet.py:65-82, third case (r=7.5e13 -> negative distance, inward)
    d, ds = oracle.distance_boundary(7.5e13, -0.3, 6.912e14, 8.64e14)
    np.testing.assert_allclose(d, -664987228972291.5, rtol=1e-15)
    assert ds == -1


def test_calculate_distance_line(oracle):
    # packets/tests/test_packet.py:88-136: static packet nu=0.4, mu=0.3, r=7.5e14, t=5.2e7
    r, mu, nu, t = 7.5e14, 0.3, 0.4, 5.2e7
    comov_nu = nu * oracle.doppler_factor(r / t, mu)
    d, err = oracle.distance_line(r, mu, nu, comov_nu, False, 0.2, t)
    assert err == 0
    np.testing.assert_allclose(d, 7.792353908000001e17, rtol=1e-15)
    d, err = oracle.distance_line(r, mu, nu, comov_nu, True, 0.2, t)
    assert d == 1e99 and err == 0
    for nu_line in (0.5, 0.6):  # MonteCarloException in the reference
        _, err = oracle.distance_line(r, mu, nu, comov_nu, False, nu_line, t)
        assert err == 1


def test_doppler_factors(oracle):
    # tardis/transport/tests/test_doppler_factor.py:9-188 (assert_almost_equal, 7 decimals, as the reference does)
    from numpy.testing import assert_almost_equal

    assert_almost_equal(oracle.doppler_factor(7.5e14 * (1 / 5.2e7), 0.3), 0.9998556693818854)
    assert_almost_equal(oracle.doppler_factor(0.0, -0.3), 1.0)
    beta = 0.2
    assert_almost_equal(oracle.doppler_factor(beta * C, 0.3), 0.94)
    assert_almost_equal(oracle.doppler_factor(beta * C, 0.3, True), 0.95938348)
    assert_almost_equal(oracle.inverse_doppler_factor(beta * C, 0.3), 1 / 0.94)
    assert_almost_equal(oracle.inverse_doppler_factor(beta * C, 0.3, True), 1.0818579)


def _two_line_model(tau, n_e, r_outer=8.64e14):
    # tests/test_transport.py:183-311: nu=4e14, lines [3.999e14, 3.998e14], t=5.2e7, seed 1963
    from tardis_b200 import synthetic as syn

    return syn.Model(
        r_inner=np.array([6.912e14]), r_outer=np.array([r_outer]), v_inner=np.array([0.0]), v_outer=np.array([0.0]),
        time_explosion=5.2e7, electron_density=np.array([n_e]), t_electrons=np.array([1e4]),
        line_list_nu=np.array([3.999e14, 3.998e14]), tau_sobolev=np.full((2, 1), tau), macro=syn.scatter_dummy_macro(),
        spectrum_frequency_grid=np.linspace(1e14, 1e15, 11), line_interaction_type="scatter")


def _one_packet():
    from tardis_b200 import synthetic as syn

    # mu chosen so that the partial-relativity lab-frame nu after the initial transform stays ~4e14
    return syn.Packets(np.array([7.5e14]), np.array([4e14]), np.array([0.3]), np.array([0.9]),
                       np.array([1963], dtype=np.int64), 1.0)


@pytest.mark.parametrize(
    "chi_over_ne, tau, r_outer, disable, first_type, first_line",
    [
        (1e-20, 0.0, 8.64e14, False, 1, None),   # BOUNDARY
        (1e-12, 0.0, 8.64e14, False, 4, None),   # ESCATTERING
        (1e-20, 100.0, 2e16, False, 2, 0),       # LINE on line 0
        (1e-12, 100.0, 2e16, True, 4, None),     # line scattering disabled -> ESCATTERING
    ],
)
def test_classic_trace_packet_structural(oracle, chi_over_ne, tau, r_outer, disable, first_type, first_line):
    """Structural assertions of test_classic_trace_packet (tests/test_transport.py:183-271):
    the first real event of the packet has the expected type."""
    model = _two_line_model(tau, n_e=1.0, r_outer=r_outer)
    res = oracle.run_oracle(model, _one_packet(), sigma_thomson=chi_over_ne, disable_line_scattering=disable,
                            n_tracked_packets=1, max_events_per_packet=4096)
    ev = res["events"][0]
    assert ev["interaction_type"][0] == 1 and ev["before_shell_id"][0] == -1  # synthetic start event
    assert ev["interaction_type"][1] == first_type
    if first_line is not None:
        assert ev["line_absorb_id"][1] == first_line
