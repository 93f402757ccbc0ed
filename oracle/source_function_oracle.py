"""CPU restatement of the formal-integral source function (second half).

TEST INFRASTRUCTURE ONLY.  numpy / scipy restatement of, paths relative to /root/reference/tardis/:
    SourceFunctionSolver.calculate_e_dot_u     spectrum/formal_integral/source_function.py:146-231
    SourceFunctionSolver.calculate_att_S_ul    spectrum/formal_integral/source_function.py:233-292
    SourceFunctionSolver.calculate_Jblue_lu    spectrum/formal_integral/source_function.py:294-334
    SourceFunctionSolver.calculate_Jred_lu     spectrum/formal_integral/source_function.py:336-358
Pinned by tests/golden/source_function_*.npz, which come from the UNMODIFIED reference class
(oracle/reference_runner.py::run_reference_source_function).

Conventions (tardis_b200.synthetic.AtomicData): one species, level index == level number; macro-atom row t belongs to line
transition_line_idx[t]; its source / destination levels are (upper, lower) of that line for emission-down (-1) and internal-down
(0) rows and (lower, upper) for internal-up (1) rows."""
from __future__ import annotations

import numpy as np

C = 2.99792458e10  # tardis/constants.py (CODATA-2010 cgs)


def transition_levels(atomic):
    """(source level, destination level) of every macro-atom row"""
    line = atomic.transition_line_idx
    up = atomic.transition_type == 1
    src = np.where(up, atomic.lower_level[line], atomic.upper_level[line])
    dst = np.where(up, atomic.upper_level[line], atomic.lower_level[line])
    return src, dst


def e_dot_u(atomic, tau_sobolev, e_dot_lu_estimator, transition_probabilities, time_of_simulation, volume, line_interaction_type):
    """source_function.py:196-231.  Returns (levels that are the upper level of some line -- ascending, as groupby gives them --,
    e_dot_u[len(levels), S])"""
    import scipy.sparse as sp
    import scipy.sparse.linalg as linalg

    norm = 1 / (time_of_simulation * volume)
    exptau = 1 - np.exp(-tau_sobolev)
    e_dot_lu = norm * exptau * e_dot_lu_estimator
    levels = np.unique(atomic.upper_level)
    pos = np.searchsorted(levels, atomic.upper_level)
    out = np.zeros((len(levels), e_dot_lu.shape[1]))
    np.add.at(out, pos, e_dot_lu)  # (pandas sums each group with Kahan compensation: equal to ~1e-16 relative)
    if line_interaction_type == "macroatom":
        n = atomic.n_levels
        internal = atomic.transition_type >= 0
        src, dst = transition_levels(atomic)
        solved = np.empty_like(out)
        for shell in range(out.shape[1]):
            q = sp.coo_matrix((transition_probabilities[internal, shell], (src[internal], dst[internal])), shape=(n, n))
            inv_n = sp.identity(n) - q
            vec = np.zeros(n)
            vec[levels] = out[:, shell]
            solved[:, shell] = linalg.spsolve(inv_n.T.tocsc(), vec)[levels]
        out = solved
    return levels, out


def solve(atomic, tau_sobolev, transition_probabilities, j_blue_estimator, e_dot_lu_estimator, time_explosion, time_of_simulation,
          volume, line_interaction_type="macroatom"):
    """SourceFunctionSolver.solve (source_function.py:27-143) on plain arrays: dict of att_S_ul, Jred_lu, Jblue_lu [L,S],
    e_dot_u [levels,S] and the level numbers of its rows"""
    levels, edu = e_dot_u(atomic, tau_sobolev, e_dot_lu_estimator, transition_probabilities, time_of_simulation, volume, line_interaction_type)
    emission = atomic.transition_type == -1
    em_line = atomic.transition_line_idx[emission]
    q_ul = transition_probabilities[emission]
    src, _ = transition_levels(atomic)
    e_rows = edu[np.searchsorted(levels, src[emission])]
    wave = atomic.wavelength_cm[em_line].reshape(-1, 1)
    att = wave * (q_ul * e_rows) * time_explosion / (4 * np.pi)  # :280
    att_s_ul = np.empty_like(tau_sobolev)
    att_s_ul[em_line] = att                                       # result.loc[line_idx]: every line has one emission row
    jblue_norm = C * time_explosion / (4 * np.pi * time_of_simulation * volume)  # :320-328
    jblue_lu = j_blue_estimator * jblue_norm
    jred_lu = jblue_lu * np.exp(-tau_sobolev) + att_s_ul          # :358
    return dict(att_S_ul=att_s_ul, Jred_lu=jred_lu, Jblue_lu=jblue_lu, e_dot_u=edu, e_dot_u_levels=levels)
