"""Load the UNMODIFIED reference hot path (TARDIS Numba code) from /root/reference.

TEST INFRASTRUCTURE ONLY.  Used by tests/golden/make_golden.py (run in the build
container, where /root/reference exists) to generate golden vectors, and by the
optional "oracle vs reference" cross-checks.  Nothing in the product
(tardis_b200/) may import this module, and nothing that runs on the GPU box may
depend on /root/reference being present.

The full `tardis` package cannot be imported here (astropy, h5py, ... are not
installed), but the Monte Carlo hot path (L0-L2 of SURVEY.md) imports fine once
a handful of heavy parents are replaced by empty namespace packages and
`astropy.units` / `tardis.constants` by tiny stand-ins.
"""
from __future__ import annotations

import os
import sys
import types

REF = os.environ.get("TARDIS_REFERENCE", "/root/reference")


class _Q:
    """Minimal stand-in for an astropy Quantity (value + no-op unit algebra)."""

    def __init__(self, value):
        self.value = value

    def to(self, *a, **k):
        return self

    @property
    def cgs(self):
        return self

    @property
    def esu(self):
        return self

    @property
    def gauss(self):
        return self

    def _v(self, o):
        return o.value if isinstance(o, _Q) else o

    def __mul__(self, o):
        return _Q(self.value * self._v(o))

    __rmul__ = __mul__

    def __truediv__(self, o):
        return _Q(self.value / self._v(o))

    def __rtruediv__(self, o):
        return _Q(self._v(o) / self.value)

    def __pow__(self, p):
        return _Q(self.value**p)

    def __float__(self):
        return float(self.value)

    # (enough of an array-valued Quantity for mc_rad_field_solver.py / planck_rad_field.py: oracle/reference_runner.py::run_reference_radfield)
    def __len__(self):
        return len(self.value)

    def __getitem__(self, k):
        return _Q(self.value[k])

    def __gt__(self, o):
        return self.value > self._v(o)

    def __ge__(self, o):
        return self.value >= self._v(o)

    def __lt__(self, o):
        return self.value < self._v(o)

    def copy(self):
        import numpy as np  # noqa: PLC0415

        return _Q(np.array(self.value, copy=True))


def available() -> bool:
    return os.path.isdir(os.path.join(REF, "tardis", "transport", "montecarlo"))


_loaded = None


def load():
    """Return a namespace with the reference's own (unmodified) hot-path objects."""
    global _loaded
    if _loaded is not None:
        return _loaded
    if not available():
        raise RuntimeError(f"reference tree not found at {REF}")

    # --- astropy stand-in -------------------------------------------------
    if "astropy" not in sys.modules:
        astropy = types.ModuleType("astropy")
        astropy.__path__ = []
        units = types.ModuleType("astropy.units")

        def _getattr(name):
            if name == "Quantity":
                return _Q
            if name.startswith("__"):
                raise AttributeError(name)
            return _Q(1.0)

        units.__getattr__ = _getattr
        astropy.units = units
        sys.modules["astropy"] = astropy
        sys.modules["astropy.units"] = units

    # --- radioactivedecay stand-in (imported by opacities/opacities.py) ----
    if "radioactivedecay" not in sys.modules:
        rd = types.ModuleType("radioactivedecay")

        class Nuclide:
            _m = {"Si-28": 27.976926535, "Fe-56": 55.9349363}

            def __init__(self, name):
                self.atomic_mass = self._m.get(name, 1.0)

        rd.Nuclide = Nuclide
        sys.modules["radioactivedecay"] = rd

    # --- namespace packages that bypass heavy __init__s ----------------------
    def ns(name, sub):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, "tardis", sub)] if sub is not None else [os.path.join(REF, "tardis")]
        sys.modules[name] = m
        return m

    ns("tardis", None)
    for name, sub in [
        ("tardis.model", "model"),
        ("tardis.model.geometry", "model/geometry"),
        ("tardis.opacities", "opacities"),
        ("tardis.plasma", "plasma"),
        ("tardis.plasma.radiation_field", "plasma/radiation_field"),
        ("tardis.transport", "transport"),
        ("tardis.io", "io"),
        ("tardis.io.logger", "io/logger"),
        ("tardis.configuration", "configuration"),
        ("tardis.util", "util"),
    ]:
        ns(name, sub)
    sys.modules["tardis.plasma.radiation_field"].DilutePlanckianRadiationField = object

    # CODATA-2010 cgs constants (astropy.constants.astropyconst13, tardis/constants.py:1)
    const = types.ModuleType("tardis.constants")
    const.c = _Q(2.99792458e10)
    const.h = _Q(6.62606957e-27)
    const.k_B = _Q(1.3806488e-16)
    const.sigma_T = _Q(6.652458734e-25)
    const.m_e = _Q(9.10938291e-28)
    const.m_p = _Q(1.672621777e-24)
    const.sigma_sb = _Q(5.670373e-5)
    const.alpha = _Q(7.2973525698e-3)
    const.e = _Q(4.80320425e-10)
    const.u = _Q(1.660538921e-24)
    const.a0 = _Q(5.2917721092e-9)
    const.Ryd = _Q(109737.31568539)
    sys.modules["tardis.constants"] = const
    sys.modules["tardis"].constants = const

    pb = types.ModuleType("tardis.transport.montecarlo.progress_bars")
    pb.update_packets_pbar = lambda *a, **k: None
    pb.refresh_packet_pbar = lambda *a, **k: None
    pb.reset_packet_pbar = lambda *a, **k: None
    pb.update_iterations_pbar = lambda *a, **k: None
    sys.modules["tardis.transport.montecarlo.progress_bars"] = pb

    R = types.SimpleNamespace()
    import tardis.transport.montecarlo as mc  # real __init__ (sets njit_dict)

    from tardis.model.geometry.radial1d_homologous import NumbaHomologousRadial1DGeometry
    from tardis.opacities.opacity_state_numba import OpacityStateNumba
    from tardis.transport.montecarlo.configuration.base import MonteCarloConfiguration
    from tardis.transport.montecarlo.modes.classic.packet_propagation import packet_propagation
    from tardis.transport.montecarlo.modes.montecarlo_transport import (
        montecarlo_transport_with_vpackets,
    )
    from tardis.transport.montecarlo.packets.packet_collections import PacketCollection
    from tardis.transport.montecarlo.packets.trackers.tracker_full_util import (
        generate_tracker_full_list,
        trackers_full_to_df,
    )
    from tardis.transport.montecarlo.packets.trackers.tracker_last_interaction_util import (
        generate_tracker_last_interaction_list,
    )

    R.mc = mc
    R.NumbaHomologousRadial1DGeometry = NumbaHomologousRadial1DGeometry
    R.OpacityStateNumba = OpacityStateNumba
    R.MonteCarloConfiguration = MonteCarloConfiguration
    R.packet_propagation = packet_propagation
    R.montecarlo_transport_with_vpackets = montecarlo_transport_with_vpackets
    R.PacketCollection = PacketCollection
    R.generate_tracker_full_list = generate_tracker_full_list
    R.trackers_full_to_df = trackers_full_to_df
    R.generate_tracker_last_interaction_list = generate_tracker_last_interaction_list
    _loaded = R
    return R
