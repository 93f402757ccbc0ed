"""Host mirror of `FormalIntegralSolver` (tardis/spectrum/formal_integral/formal_integral_solver.py:28-430) over the device
integral.

`FormalIntegralSolverB200(points, interpolate_shells, method, engine=...).solve(frequencies, simulation_state, transport_solver,
opacity_state, atomic_data, electron_densities, macro_atom_state)` takes the reference's own arguments: it runs the source
function on the device (`SourceFunctionSolverB200`) and then `tb200_formal_integral` on the tables that solve left in HBM -- the
interpolated [L, S2] tables of the reference exist only on the device.  The spectrum comes back with the attributes of the
reference's `TARDISSpectrum` that do not need astropy (`_frequency` bin edges, `frequency`, `delta_frequency`, `luminosity`,
`luminosity_density_nu`); when astropy is importable the values are Quantities, as in the reference.

The engine must hold the model of the last iteration (with the option `keep_opacity_tables = 1` for the source function) and,
by default, the line estimators of its last transport."""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from .source_function import SourceFunctionSolverB200


def _value(x):
    return getattr(x, "value", x)


class IntegrationError(Exception):
    """Same name as the reference's (spectrum/formal_integral/base.py:21)."""


def check_formal_integral_requirements(simulation_state, opacity_state, transport, raises=True):
    """spectrum/formal_integral/base.py:26-83 with the same messages; the continuum flag is read from the transport solver
    (`continuum_processes_enabled`) instead of the reference's module global."""
    import warnings

    def raise_or_return(message):
        if raises:
            raise IntegrationError(message)
        warnings.warn(message)
        return False

    for obj in (simulation_state, opacity_state, transport):
        if obj is None:
            return raise_or_return("The integrator is missing either model, opacity state or transport. Please make sure these are "
                                   "provided to the FormalIntegrator.")
    if transport.line_interaction_type not in ["downbranch", "macroatom"]:
        return raise_or_return('The FormalIntegrator currently only works for line_interaction_type == "downbranch"'
                               'and line_interaction_type == "macroatom"')
    if getattr(transport, "continuum_processes_enabled", False):
        return raise_or_return("The FormalIntegrator currently does not work for continuum interactions.")
    return True


@dataclass
class FormalIntegralSpectrum:
    """The part of TARDISSpectrum (spectrum/spectrum.py:9-54) that is plain arrays: bin edges and the luminosity per bin."""

    _frequency: object       # [n + 1] bin edges, Hz
    luminosity: object       # [n] erg / s

    @property
    def frequency(self):
        return self._frequency[:-1]

    @property
    def delta_frequency(self):
        return self._frequency[1] - self._frequency[0]

    @property
    def luminosity_density_nu(self):
        return self.luminosity / self.delta_frequency


class FormalIntegralSolverB200:
    def __init__(self, points: int, interpolate_shells: int, method: str | None = None, engine=None,
                 use_resident_estimators: bool = True) -> None:
        self.points = points
        self.interpolate_shells = interpolate_shells
        self.method = method  # "numba" / "cuda" in the reference; there is one implementation here
        self.engine = engine
        self.use_resident_estimators = use_resident_estimators
        self.timings = None
        self.source_function_state = None

    def solve(self, frequencies, simulation_state, transport_solver, opacity_state, atomic_data, electron_densities,
              macro_atom_state=None):
        if self.engine is None:
            raise ValueError("FormalIntegralSolverB200 needs the engine that holds the iteration's tables and estimators")
        check_formal_integral_requirements(simulation_state, opacity_state, transport_solver)
        if opacity_state is None or macro_atom_state is None:  # FormalIntegralSolver.setup, :78-81
            raise NotImplementedError("This functionality does not work anymore. Both opacity_state and macro_atom_state must be provided.")
        interpolate_shells = self.interpolate_shells
        if interpolate_shells == 0:  # :208-214
            interpolate_shells = max(2 * simulation_state.no_of_shells, 80)
        self.interpolate_shells = interpolate_shells
        transport_state = transport_solver.transport_state
        sf = SourceFunctionSolverB200(transport_solver.line_interaction_type, self.engine, self.use_resident_estimators)
        self.source_function_state = sf.solve(simulation_state, None, transport_state, atomic_data, macro_atom_state)
        lo = getattr(simulation_state.geometry, "v_inner_boundary_idx", 0)
        hi = getattr(simulation_state.geometry, "v_outer_boundary_idx", None)
        ne = np.asarray(_value(getattr(electron_densities, "iloc", electron_densities)[lo:hi]), dtype=np.float64)  # :354-362
        nu = np.asarray(_value(frequencies), dtype=np.float64)
        delta = nu[1] - nu[0]
        if not np.allclose(np.diff(nu), delta, atol=0, rtol=1e-12):  # :278-280
            raise AssertionError("Frequency grid must be uniform")
        res = self.engine.formal_integral(inner_temperature=float(_value(simulation_state.t_inner)), frequencies=nu, points=self.points,
                                          interpolate_shells=interpolate_shells, electron_densities=ne)
        self.timings = dict(interpolation_ms=res["interpolation_ms"], integral_ms=res["integral_ms"])
        luminosity = res["luminosity_densities"] * delta  # :281-284
        edges = np.concatenate([nu, [nu[-1] + np.diff(nu)[-1]]])  # :289-298 ("Ugly hack to convert to 'bin edges'")
        try:
            from astropy import units as u  # noqa: PLC0415

            return FormalIntegralSpectrum(u.Quantity(edges, "Hz"), u.Quantity(luminosity, "erg / s"))
        except Exception:  # astropy absent (or a stand-in without Quantity semantics): plain arrays
            return FormalIntegralSpectrum(edges, luminosity)
