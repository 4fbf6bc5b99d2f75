"""Becke partitioning: atomic weights are Becke's fuzzy-cell functions.

Counterpart of the reference's ``BeckeWPart`` (becke.py:33-120).  The reference evaluates qc-grid's
``BeckeWeights.compute_atom_weight`` per atom on the atom's own grid (O(natom^2 n_a) NumPy); here
the weights of all owner blocks come from one launch of ``hp_becke_weights`` (one thread per
point, far cells pruned against the nearest atom) on the device-resident slab.
"""

from __future__ import annotations

import json
import pathlib

import numpy as np

from . import _lib
from .core.base import WPart
from .core.logging import deflist
from .utils import ANGSTROM

__all__ = ["BeckeWPart", "becke_radii"]

_RADII = json.loads((pathlib.Path(__file__).resolve().parent / "data" / "element_radii.json").read_text())


def becke_radii(numbers):
    """Radii (bohr) as close as possible to Becke's paper: 0.35 A for hydrogen, Bragg-Slater where
    tabulated, Cordero's covalent radii otherwise (becke.py:90-102)."""
    radii = []
    for z in numbers:
        z = int(z)
        if z == 1:
            radii.append(0.35 * ANGSTROM)
            continue
        value = _RADII["radius_becke"].get(str(z), _RADII["radius_covalent"].get(str(z)))
        if value is None:
            raise ValueError(f"no Becke radius for element {z}")
        radii.append(value * ANGSTROM)
    return np.array(radii)


class BeckeWPart(WPart):
    """Becke partitioning with Becke-Lebedev grids"""

    name = "b"
    options = ["lmax", "k"]
    linear = True

    def __init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens=None, lmax=3, k=3,
                 logger=None, grid_type=1, **kwargs):  # fmt: skip
        self._k = k
        device_kw = {key: kwargs[key] for key in ("device", "comm") if key in kwargs}
        WPart.__init__(self, coordinates, numbers, pseudo_numbers, grid, moldens, spindens, lmax, logger,
                       grid_type=grid_type, **device_kw)  # fmt: skip

    k = property(lambda self: self._k)

    def _init_log_scheme(self):
        self.logger.info(" Initialized: %s" % self.__class__.__name__)
        deflist(self.logger, [(" Scheme", "Becke"), (" Switching function", "k=%i" % self._k)])

    def update_at_weights(self):
        from .core.device import stream_ptr, to_device

        if not self.local:
            raise NotImplementedError("the Becke scheme needs atomic grids (grid_type 1 or 2)")
        self.logger.info("Computing Becke weights.")
        R = becke_radii(self.numbers)
        chi = R[:, None] / R[None, :]
        u = (chi - 1) / (chi + 1)
        aab = np.clip(u / (u * u - 1), -0.45, 0.45)  # qc-grid BeckeWeights: a_ij, |a_ij| <= 0.45
        xyz = self.coordinates
        rab = np.sqrt(((xyz[:, None, :] - xyz[None, :, :]) ** 2).sum(-1))
        with np.errstate(divide="ignore"):
            inv_rab = 1.0 / rab
        s = self.slab
        dev = s.device
        _lib.call("hp_becke_weights", s.npts, s.px, s.py, s.pz, s.point_base, self.natom, s.atom_xyz,
                  s.atom_point_offsets, to_device(inv_rab, dev), to_device(aab, dev), int(self._k), s.at_w,
                  stream_ptr(dev))  # fmt: skip
        at_w = s.at_w.cpu().numpy()
        off, lo = s.atom_point_offsets_host, s.point_base
        for a in range(s.shard.atom_lo, s.shard.atom_hi):
            self.cache.dump(f"at_weights_{a}", at_w[off[a] - lo : off[a + 1] - lo])
