"""Formal integral on the device (tardis_b200/csrc/formal_integral.cuh, tb200_formal_integral,
tardis_b200/formal_integral.py).

CPU: (1) the oracle (oracle/formal_integral_oracle.{c,py}) against golden vectors computed by the reference's own
`interpolate_integrator_quantities` + `numba_formal_integral`, and against the known answers of the reference's
test_numba_formal_integral.py / test_cuda_formal_integral.py; (2) the PRODUCT's functions -- the header the CUDA kernels compile,
built for the host by tests/formal_integral_shim.cpp with the kernels' loops -- against the same goldens, against the oracle at
frequencies the goldens avoid (rays that leave the line list on the red side), and its interpolation set-up against scipy.
GPU: the kernels through the C-ABI against the goldens (tables given on the host), on the tables tb200_solve_source_function left
in HBM, and through the host mirror.
Bar: the reference's own bar between its Numba and its CUDA integrator is rtol 1e-14 on L_nu
(spectrum/formal_integral/tests/test_cuda_formal_integral.py:310); the Numba loop is fastmath, so the goldens themselves carry
that much freedom (oracle vs golden: <= 1.1e-14 on single rays).  Here: rtol 1e-13 on every intensity and luminosity density against
the goldens, 1e-14 product against oracle on L_nu (both plain IEEE; what is left is exp)."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from golden_util import GOLDEN_DIR, make_golden

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CASES = list(make_golden.FORMAL_INTEGRAL_CASES)
RTOL_GOLDEN = 1e-13
C_CGS = 2.99792458e10


def load(name):
    return make_golden.formal_integral_inputs(name), dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz")))


def oracle_solve(i, frequencies=None, points=None, interpolate_shells=None):
    from oracle import formal_integral_oracle as fio

    m = i["model"]
    return fio.solve(m.r_inner, m.r_outer, m.time_explosion, m.line_list_nu, i["inner_temperature"],
                     i["frequencies"] if frequencies is None else frequencies, i["att_S_ul"], i["Jred_lu"], i["Jblue_lu"], i["tau_sobolev"],
                     i["electron_densities"], i["points"] if points is None else points,
                     i["interpolate_shells"] if interpolate_shells is None else interpolate_shells)


def check_against_golden(got, g, rtol=RTOL_GOLDEN):
    for k in ("intensities_nu_p", "luminosity_densities"):
        a, b = np.asarray(got[k]), g[k]
        assert a.shape == b.shape, k
        assert np.array_equal(a == 0, b == 0), f"{k}: zero pattern differs"
        np.testing.assert_allclose(a, b, rtol=rtol, atol=0, err_msg=k)


# ---- (1) oracle --------------------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_oracle_matches_reference_golden(name):
    i, g = load(name)
    got = oracle_solve(i)
    check_against_golden(got, g)
    for k in ("electron_densities_interpolated", "r_inner_interpolated", "r_outer_interpolated"):
        assert np.array_equal(got[k], g[k]), k
    for k in ("att_S_ul_interpolated", "Jred_lu_interpolated", "Jblue_lu_interpolated", "tau_sobolevs_interpolated"):
        np.testing.assert_allclose(got[k].sum(axis=0), g[k + "__shell_sums"], rtol=1e-15, err_msg=k)
        np.testing.assert_allclose(got[k].sum(axis=1), g[k + "__line_sums"], rtol=1e-15, err_msg=k)


def _oracle_lib():
    from oracle import cpu_oracle

    lib = cpu_oracle.lib()
    lib.tb_oracle_fi_intersection_point.restype = C.c_double
    lib.tb_oracle_fi_intersection_point.argtypes = [C.c_double] * 3
    lib.tb_oracle_fi_black_body.restype = C.c_double
    lib.tb_oracle_fi_black_body.argtypes = [C.c_double] * 2
    lib.tb_oracle_fi_populate.restype = C.c_int64
    lib.tb_oracle_fi_populate.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    lib.tb_oracle_fi_line_search.restype = C.c_int64
    lib.tb_oracle_fi_line_search.argtypes = [C.c_void_p, C.c_double, C.c_int64]
    return lib


KAT_GEOMETRIES = [np.linspace(1, 2, 3), np.linspace(0, 1, 3)]  # test_numba_formal_integral.py:12-20 (time_explosion = 1 / c)
C_INV = 3.33564e-11


@pytest.mark.parametrize("r", KAT_GEOMETRIES)
@pytest.mark.parametrize("p", [0.0, 0.5, 1.0])
def test_oracle_kat_intersection_point(r, p):
    """test_numba_formal_integral.py:44-56."""
    lib = _oracle_lib()
    inv_t = C_CGS
    for radius in r[1:]:
        actual = lib.tb_oracle_fi_intersection_point(radius, p, inv_t)
        if p >= radius:
            assert actual == 0
        else:
            np.testing.assert_almost_equal(actual, np.sqrt(radius * radius - p * p) * C_INV * inv_t)


@pytest.mark.parametrize("r", KAT_GEOMETRIES)
@pytest.mark.parametrize("p", [0, 0.5, 1])
def test_oracle_kat_populate_photosphere(r, p):
    """test_numba_formal_integral.py:59-82: p <= r_inner[0], every shell is hit once from the inside out."""
    lib = _oracle_lib()
    r_inner, r_outer = np.ascontiguousarray(r[:-1]), np.ascontiguousarray(r[1:])
    size = len(r_outer)
    p = r_inner[0] * p
    oz, ids = np.zeros(size), np.zeros(size, dtype=np.int64)
    n = lib.tb_oracle_fi_populate(r_inner.ctypes.data, r_outer.ctypes.data, size, 1 / C_CGS, p, oz.ctypes.data, ids.ctypes.data)
    assert n == size
    assert np.array_equal(ids, np.arange(size))
    np.testing.assert_allclose(oz, 1 - np.sqrt(r_outer * r_outer - p * p), atol=1e-5)


@pytest.mark.parametrize("r", KAT_GEOMETRIES)
@pytest.mark.parametrize("p", [1e-5, 0.5, 0.99, 1])
def test_oracle_kat_populate_shells(r, p):
    """test_numba_formal_integral.py:85-128: p > r_inner[0], every crossed shell is hit twice."""
    lib = _oracle_lib()
    r_inner, r_outer = np.ascontiguousarray(r[:-1]), np.ascontiguousarray(r[1:])
    size = len(r_inner)
    p = r_inner[0] + (r_outer[-1] - r_inner[0]) * p
    idx = np.searchsorted(r_outer, p, side="right")
    offset = size - idx
    expected_n = offset * 2
    expected_ids = np.zeros(2 * size, dtype=np.int64)
    expected_ids[:expected_n] = np.abs(np.arange(0.5, expected_n, 1) - offset) - 0.5 + idx
    expected_oz = np.zeros(2 * size)
    expected_oz[0:offset] = 1 + np.sqrt(r_outer[np.arange(size, idx, -1) - 1] ** 2 - p * p)
    expected_oz[offset:expected_n] = 1 - np.sqrt(r_outer[np.arange(idx, size, 1)] ** 2 - p * p)
    oz, ids = np.zeros(2 * size), np.zeros(2 * size, dtype=np.int64)
    n = lib.tb_oracle_fi_populate(r_inner.ctypes.data, r_outer.ctypes.data, size, 1 / C_CGS, p, oz.ctypes.data, ids.ctypes.data)
    assert n == expected_n
    assert np.array_equal(ids, expected_ids)
    np.testing.assert_allclose(oz, expected_oz, atol=1e-5)


@pytest.mark.parametrize("nu,temperature", [(1e14, 1e4), (0, 1), (1, 1)])
def test_oracle_kat_black_body(nu, temperature):
    """test_cuda_formal_integral.py:34-57 (against the formula of base.py:104-120, NaN at frequency 0)."""
    got = _oracle_lib().tb_oracle_fi_black_body(nu, temperature)
    if nu == 0:
        assert np.isnan(got)
    else:
        want = 2 * 6.62606957e-27 * C_INV * C_INV * nu**3 / (np.exp(6.62606957e-27 * nu / (1.3806488e-16 * temperature)) - 1)
        np.testing.assert_allclose(got, want, rtol=1e-14)


@pytest.mark.parametrize("nu_insert", [*np.linspace(3e12, 3e16, 10), 288786721666522.1])
def test_oracle_kat_line_search(nu_insert):
    """test_cuda_formal_integral.py:205-260, on a synthetic descending list: line_search == n - searchsorted(reversed, x, "right")
    inside the list, 0 above its bluest and n below its reddest entry."""
    nu = np.ascontiguousarray(make_golden.formal_integral_inputs(CASES[0])["model"].line_list_nu)
    got = _oracle_lib().tb_oracle_fi_line_search(nu.ctypes.data, nu_insert, len(nu))
    want = 0 if nu_insert > nu[0] else len(nu) if nu_insert < nu[-1] else len(nu) - np.searchsorted(nu[::-1], nu_insert, side="right")
    assert got == want


# ---- (2) the product's header on the CPU ----------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def shim():
    out = os.path.join(HERE, "_shim", "libformal_integral_shim.so")
    src = os.path.join(HERE, "formal_integral_shim.cpp")
    hdr = os.path.join(ROOT, "tardis_b200", "csrc", "formal_integral.cuh")
    os.makedirs(os.path.dirname(out), exist_ok=True)
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.run(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", out, src], check=True)
    lib = C.CDLL(out)
    lib.fi_shim_linspace_at.restype = C.c_double
    lib.fi_shim_linspace_at.argtypes = [C.c_double, C.c_double, C.c_int, C.c_int]
    lib.fi_shim_shell_weights.argtypes = [C.c_void_p, C.c_int, C.c_double] + [C.c_void_p] * 5
    lib.fi_shim_intersection_point.restype = C.c_double
    lib.fi_shim_intersection_point.argtypes = [C.c_double] * 3
    lib.fi_shim_black_body.restype = C.c_double
    lib.fi_shim_black_body.argtypes = [C.c_double] * 2
    lib.fi_shim_count_greater.argtypes = [C.c_void_p, C.c_int, C.c_double]
    lib.fi_shim_ray_points.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    lib.fi_shim_formal_integral.argtypes = ([C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_int] + [C.c_void_p] * 11
                                            + [C.c_double, C.c_double, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int])
    return lib


def shim_weights(lib, x, x_new):
    x = np.ascontiguousarray(x, dtype=np.float64)
    lo, hi, near = C.c_int(), C.c_int(), C.c_int()
    w_lo, w_hi = C.c_double(), C.c_double()
    lib.fi_shim_shell_weights(x.ctypes.data, len(x), float(x_new), C.byref(lo), C.byref(hi), C.byref(w_lo), C.byref(w_hi), C.byref(near))
    return lo.value, hi.value, w_lo.value, w_hi.value, near.value


def shim_radii(lib, r_inner, r_outer, interpolate_shells):
    """What tb200_formal_integral does on the host: formal_integral_solver.py:208-232."""
    if interpolate_shells == 0:
        interpolate_shells = max(2 * len(r_inner), 80)
    if interpolate_shells < 0:
        return np.asarray(r_inner, dtype=np.float64), np.asarray(r_outer, dtype=np.float64)
    radius = np.array([lib.fi_shim_linspace_at(r_inner[0], r_outer[-1], interpolate_shells, k) for k in range(interpolate_shells)])
    return radius[:-1].copy(), radius[1:].copy()


def to_device_layout(table, lpad):
    """[L, S] -> [S][lpad] (the layout of tau_t and of the source function's tables in HBM)."""
    L, S = table.shape
    out = np.zeros((S, lpad))
    out[:, :L] = np.asarray(table).T
    return out


def run_shim(lib, i, frequencies=None, points=None, interpolate_shells=None, want_cells=False, warp_sweep=True):
    m = i["model"]
    frequencies = np.ascontiguousarray(i["frequencies"] if frequencies is None else frequencies, dtype=np.float64)
    points = i["points"] if points is None else points
    interpolate_shells = i["interpolate_shells"] if interpolate_shells is None else interpolate_shells
    L, S = i["tau_sobolev"].shape
    lpad = (L + 31) // 32 * 32
    r_in_i, r_out_i = shim_radii(lib, m.r_inner, m.r_outer, interpolate_shells)
    S2 = len(r_in_i)
    mid = (np.asarray(m.r_inner) + np.asarray(m.r_outer)) / 2.0
    mid_i = (r_in_i + r_out_i) / 2.0
    w = [shim_weights(lib, mid, x) for x in mid_i]
    lo, hi, near = (np.ascontiguousarray([t[k] for t in w], dtype=np.int32) for k in (0, 1, 4))
    w_lo, w_hi = (np.ascontiguousarray([t[k] for t in w], dtype=np.float64) for k in (2, 3))
    tabs = [to_device_layout(np.asarray(t), lpad) for t in (i["tau_sobolev"], i["att_S_ul"], i["Jred_lu"], i["Jblue_lu"])]
    nu_lines = np.ascontiguousarray(m.line_list_nu, dtype=np.float64)
    ne = np.ascontiguousarray(i["electron_densities"], dtype=np.float64)
    inup = np.full((len(frequencies), points), np.nan)
    lum = np.full(len(frequencies), np.nan)
    cells = np.zeros((S2, L + 2, 4)) if want_cells else None
    rc = lib.fi_shim_formal_integral(S, S2, r_in_i.ctypes.data, r_out_i.ctypes.data, float(m.time_explosion), L, lpad, nu_lines.ctypes.data,
                                     *(t.ctypes.data for t in tabs), lo.ctypes.data, hi.ctypes.data, near.ctypes.data, w_lo.ctypes.data,
                                     w_hi.ctypes.data, ne.ctypes.data, 6.652458734e-25, float(i["inner_temperature"]), len(frequencies),
                                     frequencies.ctypes.data, points, inup.ctypes.data, lum.ctypes.data,
                                     cells.ctypes.data if want_cells else None, int(warp_sweep))
    assert rc == 0
    return dict(luminosity_densities=lum, intensities_nu_p=inup, cells=cells, r_inner_interpolated=r_in_i, r_outer_interpolated=r_out_i)


@pytest.mark.parametrize("name", CASES)
def test_product_functions_match_reference_golden(shim, name):
    i, g = load(name)
    got = run_shim(shim, i, want_cells=True)
    check_against_golden(got, g)
    assert np.array_equal(got["r_inner_interpolated"], g["r_inner_interpolated"]) and np.array_equal(got["r_outer_interpolated"], g["r_outer_interpolated"])
    # the cells against the oracle's (= scipy's) interpolated tables: the linear formula is scipy's, so only exp may differ by an ulp
    o = oracle_solve(i)
    L = i["tau_sobolev"].shape[0]
    cells = got["cells"]
    assert np.array_equal(cells[:, :L, 1].T, o["att_S_ul_interpolated"])
    assert np.array_equal(cells[:, :L, 2].T, o["Jblue_lu_interpolated"])
    assert np.array_equal(cells[:, 1:L + 1, 3].T, o["Jred_lu_interpolated"])
    np.testing.assert_allclose(cells[:, :L, 0].T, np.exp(-o["tau_sobolevs_interpolated"]), rtol=4e-16)
    # flat addressing behind a shell's last line: the next shell's first line, 0.0 behind the last shell
    assert np.array_equal(cells[:-1, L, 2], o["Jblue_lu_interpolated"][0, 1:]) and cells[-1, L, 2] == 0.0
    assert np.array_equal(cells[:-1, L + 1, 3], o["Jred_lu_interpolated"][0, 1:]) and cells[-1, L + 1, 3] == 0.0
    np.testing.assert_allclose(got["luminosity_densities"], o["luminosity_densities"], rtol=1e-14)
    np.testing.assert_allclose(got["intensities_nu_p"], o["intensities_nu_p"], rtol=1e-14)
    # the warp's sweep (all lanes at the same line, gaps jumped) and a lane on its own walk the same rays
    alone = run_shim(shim, i, warp_sweep=False)
    assert np.array_equal(alone["intensities_nu_p"], got["intensities_nu_p"]) and np.array_equal(alone["luminosity_densities"], got["luminosity_densities"])


@pytest.mark.parametrize("name", CASES[:2])
def test_product_functions_match_oracle_behind_the_line_list(shim, name):
    """Frequencies the goldens avoid: rays whose window ends (or lies entirely) redward of the last line, a frequency far blueward
    of the first line, and one inside; 2 and 3 impact parameters (p = r_max alone; one ray through the photosphere)."""
    i, _ = load(name)
    nu = i["model"].line_list_nu
    freq = np.array([nu[-1] * 0.5, nu[-1] * 0.97, nu[-1], nu[-1] * 1.03, nu[len(nu) // 2], nu[0], nu[0] * 1.05, nu[0] * 3.0])
    for points in (2, 3, 17, i["points"]):
        got = run_shim(shim, i, frequencies=freq, points=points)
        o = oracle_solve(i, frequencies=freq, points=points)
        assert np.array_equal(got["intensities_nu_p"] == 0, o["intensities_nu_p"] == 0)
        np.testing.assert_allclose(got["intensities_nu_p"], o["intensities_nu_p"], rtol=1e-14)
        np.testing.assert_allclose(got["luminosity_densities"], o["luminosity_densities"], rtol=1e-14)


def test_interpolation_setup_matches_scipy(shim):
    """shell_weights against interp1d's own index / weight arithmetic: inside, outside (extrapolation), ON a grid point, and exactly
    half-way between two mid-points (kind="nearest" sends those to the left neighbour)."""
    from scipy.interpolate import interp1d

    rng = np.random.default_rng(7)
    for n in (2, 3, 7, 20):
        x = np.sort(rng.random(n)) + np.arange(n)
        y = rng.random((5, n))
        probes = np.concatenate([rng.random(40) * (n + 2) - 1, x, x[1:] / 2.0 + x[:-1] / 2.0, [x[0] - 3.0, x[-1] + 3.0]])
        lin = interp1d(x, y, fill_value="extrapolate")(probes)
        near = interp1d(x, y, fill_value="extrapolate", kind="nearest")(probes)
        for k, xn in enumerate(probes):
            lo, hi, w_lo, w_hi, nearest = shim_weights(shim, x, xn)
            assert np.array_equal(w_hi * y[:, hi] + w_lo * y[:, lo], lin[:, k])
            assert np.array_equal(y[:, nearest], near[:, k])


def test_linspace_matches_numpy(shim):
    for a, b, n in ((1.2355e15, 2.2464e15, 80), (0.0, 1.0, 2), (3.0, 7.0, 41), (1e15, 1.0000001e15, 13)):
        got = np.array([shim.fi_shim_linspace_at(a, b, n, k) for k in range(n)])
        assert np.array_equal(got, np.linspace(a, b, n))


@pytest.mark.parametrize("r", KAT_GEOMETRIES)
def test_product_ray_geometry_matches_oracle(shim, r):
    """Ray::point (intersection points recomputed per segment) against populate_intersection_points, every impact parameter of a
    small grid; plus the helpers' known answers."""
    lib = _oracle_lib()
    r_inner, r_outer = np.ascontiguousarray(r[:-1]), np.ascontiguousarray(r[1:])
    size, n_p, t = len(r_inner), 23, 1 / C_CGS
    for p_idx in range(1, n_p):
        p = p_idx * r_outer[-1] / (n_p - 1)
        oz, ids = np.zeros(2 * size), np.zeros(2 * size, dtype=np.int64)
        n = lib.tb_oracle_fi_populate(r_inner.ctypes.data, r_outer.ctypes.data, size, t, p, oz.ctypes.data, ids.ctypes.data)
        pz, pids = np.zeros(2 * size), np.zeros(2 * size, dtype=np.int32)
        m = shim.fi_shim_ray_points(size, r_inner.ctypes.data, r_outer.ctypes.data, t, p_idx, n_p, pz.ctypes.data, pids.ctypes.data)
        assert m == n
        assert np.array_equal(pz[:n], oz[:n]) and np.array_equal(pids[:n], ids[:n])
    assert shim.fi_shim_intersection_point(2.0, 1.0, C_CGS) == lib.tb_oracle_fi_intersection_point(2.0, 1.0, C_CGS)
    assert shim.fi_shim_intersection_point(1.0, 1.0, C_CGS) == 0.0
    assert np.isnan(shim.fi_shim_black_body(0.0, 1.0))
    assert shim.fi_shim_black_body(1e14, 1e4) == lib.tb_oracle_fi_black_body(1e14, 1e4)
    nu = np.ascontiguousarray(make_golden.formal_integral_inputs(CASES[0])["model"].line_list_nu)
    for x in [*np.linspace(3e12, 3e16, 10), nu[0], nu[-1], nu[17]]:
        assert shim.fi_shim_count_greater(nu.ctypes.data, len(nu), x) == lib.tb_oracle_fi_line_search(nu.ctypes.data, x, len(nu))



# ---- at the size bench.py runs it: 5e5 lines, 20 -> 79 shells, 1000 impact parameters ---------------------------------------------
def load_bench_shape():
    return make_golden.formal_integral_bench_shape_inputs(), dict(np.load(os.path.join(GOLDEN_DIR, "formal_integral_bench_shape.npz")))


def test_oracle_matches_reference_at_the_bench_shape():
    """The golden is the UNMODIFIED reference's output at full size (scripts/reference_formal_integral_rate.py --golden /
    make_golden.py --case formal_integral_bench_shape): windows of ~18 000 lines per ray, 998 rays per frequency."""
    i, g = load_bench_shape()
    assert np.array_equal(i["frequencies"], g["frequencies"]), "synthetic inputs drifted from the golden"
    got = oracle_solve(i)
    check_against_golden(got, g)
    assert got["tau_sobolevs_interpolated"].shape == (500_000, 79)


def test_product_functions_match_reference_at_the_bench_shape(shim):
    """The kernels' arithmetic (the product header on the CPU, the warp sweep lane by lane) at full size against the reference itself:
    four of the golden's frequencies, all 1000 impact parameters."""
    i, g = load_bench_shape()
    pick = np.array([0, 5, 10, 15])
    got = run_shim(shim, i, frequencies=i["frequencies"][pick])
    check_against_golden(got, {k: g[k][pick] for k in ("intensities_nu_p", "luminosity_densities")})

# ---- (3) the kernels through the C-ABI -------------------------------------------------------------------------------------------
def engine_for(i):
    """An engine holding the model the case's tables belong to (macro-atom metadata of the source-function case included)."""
    from tardis_b200.engine import Engine

    sf = make_golden.source_function_inputs(i["source_function_case"])
    a, m = sf["atomic"], i["model"]
    eng = Engine(0)
    eng.set_option("keep_opacity_tables", 1)
    eng.set_model(r_inner=m.r_inner, r_outer=m.r_outer, time_explosion=m.time_explosion, electron_density=m.electron_density,
                  line_list_nu=m.line_list_nu, tau_sobolev=i["tau_sobolev"], line_interaction_type=sf["mode"],
                  transition_probabilities=sf["transition_probabilities"], line2macro_level_upper=a.line2macro_level_upper,
                  macro_block_edge_index=a.macro_block_edge_index, transition_type=a.transition_type,
                  destination_level_id=a.destination_level_id, transition_line_id=a.transition_line_idx,
                  spectrum_frequency_grid=m.spectrum_frequency_grid)
    return eng, sf


def engine_integral(eng, i, frequencies=None, points=None, interpolate_shells=None, tables="host"):
    return eng.formal_integral(inner_temperature=i["inner_temperature"], frequencies=i["frequencies"] if frequencies is None else frequencies,
                               points=i["points"] if points is None else points,
                               interpolate_shells=i["interpolate_shells"] if interpolate_shells is None else interpolate_shells,
                               tables=(i["att_S_ul"], i["Jred_lu"], i["Jblue_lu"]) if tables == "host" else None, want_intensities=True)


@pytest.mark.gpu
@pytest.mark.parametrize("name", CASES)
def test_engine_matches_reference_golden(name):
    i, g = load(name)
    eng, _ = engine_for(i)
    got = engine_integral(eng, i)
    check_against_golden(got, g)
    o = oracle_solve(i)
    np.testing.assert_allclose(got["luminosity_densities"], o["luminosity_densities"], rtol=1e-14)
    np.testing.assert_allclose(got["intensities_nu_p"], o["intensities_nu_p"], rtol=1e-14)
    assert got["integral_ms"] > 0.0
    eng.close()


@pytest.mark.gpu
def test_engine_matches_oracle_behind_the_line_list_and_on_tiny_grids():
    i, _ = load(CASES[0])
    eng, _ = engine_for(i)
    nu = i["model"].line_list_nu
    freq = np.array([nu[-1] * 0.5, nu[-1] * 0.97, nu[-1], nu[-1] * 1.03, nu[len(nu) // 2], nu[0], nu[0] * 1.05, nu[0] * 3.0])
    for points, shells in ((2, 20), (3, -1), (17, 5), (33, 2), (i["points"], 0), (97, 150)):
        got = engine_integral(eng, i, frequencies=freq, points=points, interpolate_shells=shells)
        o = oracle_solve(i, frequencies=freq, points=points, interpolate_shells=shells)
        assert np.array_equal(got["intensities_nu_p"] == 0, o["intensities_nu_p"] == 0), (points, shells)
        np.testing.assert_allclose(got["intensities_nu_p"], o["intensities_nu_p"], rtol=1e-14, err_msg=str((points, shells)))
        np.testing.assert_allclose(got["luminosity_densities"], o["luminosity_densities"], rtol=1e-14, err_msg=str((points, shells)))
    # no frequency at all, and one frequency
    assert engine_integral(eng, i, frequencies=np.zeros(0))["luminosity_densities"].shape == (0,)
    one = engine_integral(eng, i, frequencies=freq[4:5])
    np.testing.assert_allclose(one["luminosity_densities"], oracle_solve(i, frequencies=freq[4:5])["luminosity_densities"], rtol=1e-14)
    eng.close()


@pytest.mark.gpu
def test_engine_error_conditions():
    from tardis_b200.engine import Engine, EngineError

    i, _ = load(CASES[0])
    eng, _ = engine_for(i)
    with pytest.raises(EngineError, match="source function"):
        engine_integral(eng, i, tables=None)  # nothing resident yet
    with pytest.raises(EngineError, match="n_impact_parameters"):
        engine_integral(eng, i, points=1)
    with pytest.raises(EngineError, match="interpolate_shells"):
        engine_integral(eng, i, interpolate_shells=1)
    with pytest.raises(ValueError):
        eng.formal_integral(inner_temperature=1e4, frequencies=i["frequencies"], points=8, tables=(i["att_S_ul"][:-1], i["Jred_lu"], i["Jblue_lu"]))
    eng.close()
    m = i["model"]
    eng = Engine(0)  # line_interaction_type scatter: check_formal_integral_requirements refuses (base.py:62-70)
    eng.set_model(r_inner=m.r_inner, r_outer=m.r_outer, time_explosion=m.time_explosion, electron_density=m.electron_density,
                  line_list_nu=m.line_list_nu, tau_sobolev=i["tau_sobolev"], line_interaction_type="scatter",
                  spectrum_frequency_grid=m.spectrum_frequency_grid)
    with pytest.raises(EngineError, match="downbranch and macroatom"):
        engine_integral(eng, i)
    eng.close()


@pytest.mark.gpu
def test_formal_integral_of_resident_tables_and_host_mirror():
    """The real sequence after the last iteration: transport leaves the line estimators in HBM, the source function reads them and
    leaves att_S_ul / Jred_lu / Jblue_lu there, the integral reads those.  Then the same through the host mirror with the
    reference's argument objects.  Comparator: the oracles (source function + formal integral) on the downloaded estimators."""
    import types

    import pandas as pd

    from oracle import formal_integral_oracle as fio
    from oracle import opacity_oracle
    from oracle import source_function_oracle as sfo
    from tardis_b200 import synthetic as syn
    from tardis_b200.engine import Engine
    from tardis_b200.formal_integral import FormalIntegralSolverB200, IntegrationError

    S, L, n_levels = 8, 4000, 300
    model = syn.make_model(S, L, "macroatom", mu_tau=-4.0, seed=71)
    atomic = syn.make_atomic_data(model.line_list_nu, n_levels, "macroatom", seed=72, nlte_fraction=0.0)
    plasma = syn.make_plasma_state(atomic, S, model.time_explosion, seed=73, zero_fraction=0.0, inversion_fraction=0.0, noise=0.0)
    plasma.level_number_density *= 1e-9
    tables = opacity_oracle.build(atomic, plasma)
    volume = 4.0 / 3.0 * np.pi * (model.r_outer**3 - model.r_inner**3)
    t_exp, t_sim, t_inner, points = float(model.time_explosion), 2.0e5, 1.1e4, 120
    eng = Engine(0)
    eng.set_option("keep_opacity_tables", 1)
    eng.set_model(r_inner=model.r_inner, r_outer=model.r_outer, time_explosion=t_exp, electron_density=model.electron_density,
                  line_list_nu=model.line_list_nu, tau_sobolev=tables["tau_sobolev"], line_interaction_type="macroatom",
                  transition_probabilities=tables["transition_probabilities"], line2macro_level_upper=atomic.line2macro_level_upper,
                  macro_block_edge_index=atomic.macro_block_edge_index, transition_type=atomic.transition_type,
                  destination_level_id=atomic.destination_level_id, transition_line_id=atomic.transition_line_idx,
                  spectrum_frequency_grid=model.spectrum_frequency_grid)
    res = eng.run_packets(syn.make_packets(50000, model.r_inner[0], base_seed=13))
    sf = eng.solve_source_function(time_explosion=t_exp, time_of_simulation=t_sim, volume=volume, wavelength_cm=atomic.wavelength_cm,
                                   lines_lower_level_idx=atomic.lower_level, lines_upper_level_idx=atomic.upper_level, n_levels=n_levels)
    nu = model.line_list_nu
    z_max = model.r_outer[-1] / t_exp / C_CGS
    freq = np.linspace(nu[-1] / (1 - z_max) * 1.001, nu[0] * 1.02, 200)
    got = eng.formal_integral(inner_temperature=t_inner, frequencies=freq, points=points, interpolate_shells=0, want_intensities=True)
    # (a) the integral alone: oracle on the tables the device solve returned
    want = fio.solve(model.r_inner, model.r_outer, t_exp, nu, t_inner, freq, sf["att_S_ul"], sf["Jred_lu"], sf["Jblue_lu"], tables["tau_sobolev"],
                     model.electron_density, points, 0)
    assert np.count_nonzero(want["luminosity_densities"]) == len(freq)
    np.testing.assert_allclose(got["intensities_nu_p"], want["intensities_nu_p"], rtol=1e-14)
    np.testing.assert_allclose(got["luminosity_densities"], want["luminosity_densities"], rtol=1e-14)
    # (b) the chain: both oracles from the downloaded estimators (the source function's bar: tests/test_source_function.py)
    sfw = sfo.solve(atomic, tables["tau_sobolev"], tables["transition_probabilities"], res["j_blue"], res["edotlu"], t_exp, t_sim, volume, "macroatom")
    chain = fio.solve(model.r_inner, model.r_outer, t_exp, nu, t_inner, freq, sfw["att_S_ul"], sfw["Jred_lu"], sfw["Jblue_lu"], tables["tau_sobolev"],
                      model.electron_density, points, 0)
    np.testing.assert_allclose(got["luminosity_densities"], chain["luminosity_densities"], rtol=1e-9)
    # (c) host mirror with the reference's argument objects
    ns = types.SimpleNamespace
    lines = pd.DataFrame({"line_id": np.arange(L), "wavelength_cm": atomic.wavelength_cm},
                         index=pd.MultiIndex.from_arrays([np.full(L, 14), np.full(L, 1), atomic.lower_level, atomic.upper_level],
                                                         names=["atomic_number", "ion_number", "level_number_lower", "level_number_upper"]))
    refs = pd.Series(np.arange(n_levels), index=pd.MultiIndex.from_arrays([np.full(n_levels, 14), np.full(n_levels, 1), np.arange(n_levels)],
                                                                          names=["atomic_number", "ion_number", "level_number"]))
    sim_state = ns(geometry=ns(v_inner_boundary_idx=0, v_outer_boundary_idx=S), no_of_shells=S, volume=volume, time_explosion=t_exp, t_inner=t_inner)
    transport_state = ns(estimators_line=ns(mean_intensity_blueward=res["j_blue"], energy_deposition_line_rate=res["edotlu"]),
                         packet_collection=ns(time_of_simulation=t_sim))
    transport_solver = ns(line_interaction_type="macroatom", transport_state=transport_state, continuum_processes_enabled=False)
    solver = FormalIntegralSolverB200(points, 0, "cuda", engine=eng)
    spectrum = solver.solve(freq, sim_state, transport_solver, ns(tau_sobolev=tables["tau_sobolev"]), ns(lines=lines), pd.Series(model.electron_density),
                            ns(references_index=refs))
    assert solver.interpolate_shells == 80
    lum = np.asarray(getattr(spectrum.luminosity, "value", spectrum.luminosity))
    edges = np.asarray(getattr(spectrum._frequency, "value", spectrum._frequency))
    assert np.array_equal(lum, got["luminosity_densities"] * (freq[1] - freq[0]))
    assert edges.shape == (len(freq) + 1,) and np.array_equal(edges[:-1], freq)
    with pytest.raises(IntegrationError):
        solver.solve(freq, sim_state, ns(line_interaction_type="scatter", transport_state=transport_state), ns(), ns(lines=lines),
                     pd.Series(model.electron_density), ns(references_index=refs))
    with pytest.raises(IntegrationError):
        solver.solve(freq, sim_state, ns(line_interaction_type="macroatom", transport_state=transport_state, continuum_processes_enabled=True), ns(),
                     ns(lines=lines), pd.Series(model.electron_density), ns(references_index=refs))
    # a new model invalidates the resident source function
    eng.set_model(r_inner=model.r_inner, r_outer=model.r_outer, time_explosion=t_exp, electron_density=model.electron_density,
                  line_list_nu=model.line_list_nu, tau_sobolev=tables["tau_sobolev"], line_interaction_type="macroatom",
                  transition_probabilities=tables["transition_probabilities"], line2macro_level_upper=atomic.line2macro_level_upper,
                  macro_block_edge_index=atomic.macro_block_edge_index, transition_type=atomic.transition_type,
                  destination_level_id=atomic.destination_level_id, transition_line_id=atomic.transition_line_idx,
                  spectrum_frequency_grid=model.spectrum_frequency_grid)
    from tardis_b200.engine import EngineError

    with pytest.raises(EngineError, match="source function"):
        eng.formal_integral(inner_temperature=t_inner, frequencies=freq, points=points)
    eng.close()


@pytest.mark.gpu
def test_engine_matches_oracle_on_a_wide_grid():
    """More rays than one warp-block per frequency, windows of many hundred lines, 79 integrator shells from 20: every lane pattern of
    the sweep (lanes that start late, finish early, rays through the photosphere next to rays that miss it)."""
    from oracle import formal_integral_oracle as fio
    from tardis_b200 import synthetic as syn
    from tardis_b200.engine import Engine

    S, L, points = 20, 30000, 333
    model = syn.make_model(S, L, "downbranch", mu_tau=-3.0, seed=81)
    rng = np.random.default_rng(82)
    att = rng.random((L, S)) * 1e-6
    jblue = rng.random((L, S)) * 1e-5
    jred = jblue * np.exp(-np.asarray(model.tau_sobolev)) + att
    eng = Engine(0)
    m = model.macro
    eng.set_model(r_inner=model.r_inner, r_outer=model.r_outer, time_explosion=model.time_explosion, electron_density=model.electron_density,
                  line_list_nu=model.line_list_nu, tau_sobolev=model.tau_sobolev, line_interaction_type="downbranch",
                  transition_probabilities=m.transition_probabilities, line2macro_level_upper=m.line2macro_level_upper,
                  macro_block_edge_index=m.macro_block_edge_index, transition_type=m.transition_type,
                  destination_level_id=m.destination_level_id, transition_line_id=m.transition_line_id,
                  spectrum_frequency_grid=model.spectrum_frequency_grid)
    nu = model.line_list_nu
    freq = np.linspace(nu[-1] * 1.2, nu[0] * 0.9, 48)
    got = eng.formal_integral(inner_temperature=1e4, frequencies=freq, points=points, interpolate_shells=0, tables=(att, jred, jblue),
                              want_intensities=True)
    want = fio.solve(model.r_inner, model.r_outer, float(model.time_explosion), nu, 1e4, freq, att, jred, jblue, model.tau_sobolev,
                     model.electron_density, points, 0)
    np.testing.assert_allclose(got["intensities_nu_p"], want["intensities_nu_p"], rtol=1e-14)
    np.testing.assert_allclose(got["luminosity_densities"], want["luminosity_densities"], rtol=1e-14)
    eng.close()
