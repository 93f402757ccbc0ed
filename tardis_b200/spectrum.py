"""Host mirror of `SpectrumSolver` (tardis/spectrum/base.py:14-202) and of the plain-array part of `TARDISSpectrum`
(tardis/spectrum/spectrum.py:9-120) over what the engine already summed on the device ("fused
emitted / reabsorbed spectrum histogram ... and the hand-off to the spectrum solver").

The reference's solver histograms the per-packet arrays on the host every time a spectrum is asked for
(`np.histogram(emitted_packet_nu, weights=emitted_packet_luminosity, bins=grid)`, spectrum/base.py:139-159).  Here the transport
kernels' epilogue filled those histograms while the packets finished; `SpectrumSolverB200` takes them from the transport state
(`MonteCarloTransportState.montecarlo_emitted_luminosity` / `montecarlo_reabsorbed_luminosity`) and falls back to the reference's
formula on the per-packet arrays only when a state carries no fused sums (a state the reference produced).  The integrated spectrum
comes from `FormalIntegralSolverB200` (tardis_b200/formal_integral.py) with the reference's call (spectrum/base.py:97-105).

Same constructor, `setup_optional_spectra`, `solve`, properties and warnings as the reference; values are astropy Quantities when
astropy is importable, plain arrays otherwise."""
from __future__ import annotations

import warnings
from dataclasses import dataclass

import numpy as np

from .formal_integral import IntegrationError

C_ANGSTROM_PER_S = 2.99792458e18  # const.c in Angstrom / s (CODATA-2010, as tardis.constants)


def _value(x):
    return getattr(x, "value", x)


def _quantity(value, unit):
    try:
        from astropy import units as u  # noqa: PLC0415

        return u.Quantity(value, unit)
    except Exception:  # astropy absent: plain arrays
        return value


@dataclass
class TARDISSpectrumB200:
    """TARDISSpectrum(_frequency, luminosity) (spectrum/spectrum.py:9-54): bin edges [n + 1] in Hz, luminosity per bin [n] in erg/s,
    and the quantities its constructor derives from them."""

    _frequency: object
    luminosity: object

    def __post_init__(self):
        if np.shape(self._frequency)[0] != np.shape(self.luminosity)[0] + 1:  # spectrum.py:34-38
            raise ValueError("shape of '_frequency' and 'luminosity' are not compatible"
                             f": '{np.shape(self._frequency)[0]}' and '{np.shape(self.luminosity)[0]}'")

    @property
    def frequency(self):
        return self._frequency[:-1]

    @property
    def delta_frequency(self):
        return self._frequency[1] - self._frequency[0]

    @property
    def wavelength(self):
        """Angstrom (frequency.to("angstrom", u.spectral()), spectrum.py:46)"""
        return _quantity(C_ANGSTROM_PER_S / np.asarray(_value(self.frequency), dtype=np.float64), "angstrom")

    @property
    def luminosity_density_nu(self):
        return self.luminosity / self.delta_frequency

    @property
    def luminosity_density_lambda(self):
        """f_nu_to_f_lambda (spectrum.py:96-97): L_nu nu^2 / c, erg / s / Angstrom"""
        nu = np.asarray(_value(self.frequency), dtype=np.float64)
        l_nu = np.asarray(_value(self.luminosity), dtype=np.float64) / float(_value(self.delta_frequency))
        return _quantity(l_nu * nu ** 2 / C_ANGSTROM_PER_S, "erg / (s angstrom)")


class SpectrumSolverB200:
    hdf_properties = ["montecarlo_virtual_luminosity", "spectrum_real_packets", "spectrum_virtual_packets",
                      "spectrum_real_packets_reabsorbed", "spectrum_integrated"]
    hdf_name = "spectrum"

    def __init__(self, transport_state, spectrum_frequency_grid, integrator_settings):
        self.transport_state = transport_state
        self.spectrum_frequency_grid = spectrum_frequency_grid
        self._montecarlo_virtual_luminosity = _quantity(np.zeros(np.shape(_value(spectrum_frequency_grid))), "erg / s")
        self._integrator = None
        self.integrator_settings = integrator_settings
        self._spectrum_integrated = None
        self.simulation_state = self.opacity_state = self.transport = self.plasma = self.macro_atom_state = None

    def setup_optional_spectra(self, transport_state, virtual_packet_luminosity=None, integrator=None, simulation_state=None,
                               transport=None, plasma=None, opacity_state=None, macro_atom_state=None):
        """spectrum/base.py:37-63.  `virtual_packet_luminosity` is the unnormalised virtual-packet energy histogram
        (`v_packets_energy_hist`, what `MCTransportSolverB200.run` returns); `integrator` a `FormalIntegralSolverB200`."""
        self.transport_state = transport_state
        if virtual_packet_luminosity is not None:
            _value(self._montecarlo_virtual_luminosity)[:] = _value(virtual_packet_luminosity)
        self._integrator = integrator
        self._spectrum_integrated = None  # a new set-up belongs to a new last iteration
        self.simulation_state = simulation_state
        self.opacity_state = opacity_state
        self.transport = transport
        self.plasma = plasma
        self.macro_atom_state = macro_atom_state

    # ---- the three Monte Carlo luminosity histograms ----
    def _histogram(self, fused_name, nu_name, lum_name):
        fused = getattr(self.transport_state, fused_name, None)  # MonteCarloTransportState of tardis_b200: summed in the kernel epilogue
        if fused is not None:
            grid = np.asarray(_value(self.spectrum_frequency_grid))
            if np.shape(_value(fused))[0] == len(grid) - 1:
                return _quantity(np.asarray(_value(fused), dtype=np.float64), "erg / s")
        nu = np.asarray(_value(getattr(self.transport_state, nu_name)))  # the reference's formula (spectrum/base.py:139-159)
        lum = np.asarray(_value(getattr(self.transport_state, lum_name)))
        return _quantity(np.histogram(nu, weights=lum, bins=np.asarray(_value(self.spectrum_frequency_grid)))[0], "erg / s")

    @property
    def montecarlo_reabsorbed_luminosity(self):
        return self._histogram("montecarlo_reabsorbed_luminosity", "reabsorbed_packet_nu", "reabsorbed_packet_luminosity")

    @property
    def montecarlo_emitted_luminosity(self):
        return self._histogram("montecarlo_emitted_luminosity", "emitted_packet_nu", "emitted_packet_luminosity")

    @property
    def montecarlo_virtual_luminosity(self):
        return self._montecarlo_virtual_luminosity[:-1] / float(_value(self.transport_state.time_of_simulation))  # :162-166

    # ---- spectra ----
    @property
    def spectrum_real_packets(self):
        return TARDISSpectrumB200(self.spectrum_frequency_grid, self.montecarlo_emitted_luminosity)

    @property
    def spectrum_real_packets_reabsorbed(self):
        return TARDISSpectrumB200(self.spectrum_frequency_grid, self.montecarlo_reabsorbed_luminosity)

    @property
    def spectrum_virtual_packets(self):
        if np.all(np.asarray(_value(self.montecarlo_virtual_luminosity)) == 0):
            warnings.warn("SpectrumSolver.spectrum_virtual_packets is zero. Please run the montecarlo simulation with "
                          "no_of_virtual_packets > 0", UserWarning)
        return TARDISSpectrumB200(self.spectrum_frequency_grid, self.montecarlo_virtual_luminosity)

    @property
    def integrator(self):
        if self._integrator is None:
            warnings.warn("SpectrumSolver.integrator: The FormalIntegrator is not yet available."
                          "Please run the montecarlo simulation at least once.", UserWarning)
        if getattr(self.transport_state, "enable_full_relativity", False):
            raise NotImplementedError("The FormalIntegrator is not yet implemented for the full relativity mode. "
                                      "Please run with config option enable_full_relativity: False.")
        return self._integrator

    @property
    def spectrum_integrated(self):
        if self._spectrum_integrated is None and self.integrator is not None:
            try:
                self._spectrum_integrated = self.integrator.solve(
                    self.spectrum_frequency_grid[:-1], self.simulation_state, self.transport, self.opacity_state, self.plasma.atomic_data,
                    self.plasma.electron_densities, self.macro_atom_state)
            except IntegrationError:  # spectrum/base.py:106-119: an empty spectrum, with the reference's warning
                warnings.warn("The FormalIntegrator is not yet implemented for the full relativity mode or continuum processes. "
                              "Please run with config option enable_full_relativity: False and continuum_processes_enabled: False "
                              "This RETURNS AN EMPTY SPECTRUM!", UserWarning)
                self._spectrum_integrated = TARDISSpectrumB200(_quantity(np.array([np.nan, np.nan]), "Hz"), _quantity(np.array([np.nan]), "erg / s"))
        return self._spectrum_integrated

    def solve(self, transport_state):
        """-> (real, virtual, integrated) spectra (spectrum/base.py:168-187)"""
        self.transport_state = transport_state
        return (self.spectrum_real_packets, self.spectrum_virtual_packets, self.spectrum_integrated)

    @classmethod
    def from_config(cls, config):
        """spectrum/base.py:189-202: num + 1 edges from spectrum.stop to spectrum.start, converted to Hz with u.spectral() (the same
        grid `MCTransportSolverB200.from_config` builds)."""
        from .montecarlo import _to_hz  # noqa: PLC0415

        grid = np.linspace(_to_hz(config.spectrum.stop), _to_hz(config.spectrum.start), num=config.spectrum.num + 1)
        return cls(transport_state=None, spectrum_frequency_grid=_quantity(grid, "Hz"), integrator_settings=config.spectrum.integrated)
