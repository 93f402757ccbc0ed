"""Host mirror of `SourceFunctionSolver` (tardis/spectrum/formal_integral/source_function.py:16-143) over the device solve.

`SourceFunctionSolverB200(line_interaction_type, engine).solve(sim_state, opacity_state_numba, transport_state, atomic_data,
macro_atom_state)` takes the reference's own arguments and returns an object with the reference's attributes (`att_S_ul`,
`Jred_lu`, `Jblue_lu`, `e_dot_u`).  The tables and the line estimators are read where they lie in HBM: the engine must hold the
model of the last iteration with the option `keep_opacity_tables = 1` (the normalised transition probabilities), and by default
the estimators of its last transport; the arrays of `transport_state` / `opacity_state_numba` are not uploaded again."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

C_CGS = 2.99792458e10  # tardis/constants.py (CODATA-2010): const.c.cgs


@dataclass
class SourceFunctionState:
    """Same fields as the reference's SourceFunctionState (source_function.py:361-400)."""

    att_S_ul: np.ndarray
    Jred_lu: np.ndarray
    Jblue_lu: np.ndarray
    e_dot_u: object  # pandas.DataFrame indexed like the reference's when pandas is importable, else (levels, ndarray)


def _value(x):
    return getattr(x, "value", x)


class SourceFunctionSolverB200:
    def __init__(self, line_interaction_type: str, engine=None, use_resident_estimators: bool = True) -> None:
        if line_interaction_type not in ("downbranch", "macroatom"):
            # FormalIntegralSolver.setup refuses anything else (spectrum/formal_integral/base.py: check)
            raise ValueError("The FormalIntegrator currently only works for line_interaction_type = downbranch and macroatom")
        self.line_interaction_type = line_interaction_type
        self.engine = engine
        self.use_resident_estimators = use_resident_estimators
        self.iterations = None

    def solve(self, sim_state, opacity_state_numba, transport_state, atomic_data, macro_atom_state) -> SourceFunctionState:
        if self.engine is None:
            raise ValueError("SourceFunctionSolverB200 needs the engine that holds the iteration's tables and estimators")
        lo_idx = getattr(sim_state.geometry, "v_inner_boundary_idx", 0)
        hi_idx = getattr(sim_state.geometry, "v_outer_boundary_idx", None)
        volume = np.asarray(_value(sim_state.volume), dtype=np.float64)
        time_explosion = float(_value(sim_state.time_explosion))
        time_of_simulation = float(_value(transport_state.packet_collection.time_of_simulation))
        lines = atomic_data.lines
        references = macro_atom_state.references_index
        # level index of a line's lower / upper level = row of MacroAtomState.references_index (source_function.py:204)
        idx = lines.index
        names = list(idx.names)
        lower_key = idx.droplevel("level_number_upper") if "level_number_upper" in names else None
        upper_key = idx.droplevel("level_number_lower") if "level_number_lower" in names else None
        lower = np.asarray(references.loc[lower_key]).ravel().astype(np.int64)
        upper = np.asarray(references.loc[upper_key]).ravel().astype(np.int64)
        estimators = None
        if not self.use_resident_estimators:
            est = transport_state.estimators_line
            estimators = (np.asarray(est.mean_intensity_blueward)[:, lo_idx:hi_idx], np.asarray(est.energy_deposition_line_rate)[:, lo_idx:hi_idx])
        res = self.engine.solve_source_function(
            time_explosion=time_explosion, time_of_simulation=time_of_simulation, volume=volume,
            wavelength_cm=np.asarray(lines.wavelength_cm, dtype=np.float64), lines_lower_level_idx=lower, lines_upper_level_idx=upper,
            n_levels=len(references), estimators=estimators, c=C_CGS)
        self.iterations = res["iterations"]
        levels = np.unique(upper)  # e_dot_u.index: the groups of the reference's group-by, ascending (source_function.py:198)
        e_dot_u = res["e_dot_u"][levels]
        try:
            import pandas as pd

            index = references.index[levels] if hasattr(references, "index") else pd.Index(levels)
            e_dot_u = pd.DataFrame(e_dot_u, index=index, columns=range(e_dot_u.shape[1]))
            e_dot_u.index.names = ["atomic_number", "ion_number", "source_level_number"][: e_dot_u.index.nlevels]
        except ImportError:  # pragma: no cover
            e_dot_u = (levels, e_dot_u)
        return SourceFunctionState(res["att_S_ul"], res["Jred_lu"], res["Jblue_lu"], e_dot_u)
