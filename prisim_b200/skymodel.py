"""Minimal ``SkyModel`` container with the attributes the hot path reads.

The reference takes an ``astroutils.catalog.SkyModel`` (un-vendored dependency) and touches only
``location``, ``epoch``, ``src_shape`` and ``generate_spectrum`` inside ``observe``
(interferometry.py:6171-6174, :6249, :6258-6267); run_prisim builds it from ``init_parms`` with
``spec_type='func'`` and power-law ``spec_parms`` (scripts/run_prisim.py:1629-1641).  This class
keeps those names so catalogue-building code ports unchanged; any duck-typed object with the same
attributes is accepted by ``InterferometerArray.observe``.
"""
from __future__ import annotations

import numpy as NP


class SkyModel(object):
    """Attributes
    name        [nsrc] source names (optional)
    frequency   [1 or nchan] Hz (reference frequency axis of the catalogue)
    location    [nsrc,2] degrees (RA/Dec, HA/Dec or Alt/Az according to `coords`)
    coords      'radec' | 'hadec' | 'altaz'
    epoch       e.g. 'J2000'
    spec_type   'func' (power law) | 'spectrum' (tabulated)
    spec_parms  {'name', 'power-law-index', 'freq-ref', 'flux-scale', 'flux-offset', ...}
    spectrum    [nsrc, nfreq] for spec_type == 'spectrum'
    src_shape   None or [nsrc,3] (major FWHM deg, minor FWHM deg, position angle deg)
    """

    def __init__(self, init_file=None, init_parms=None):
        if init_file is not None:
            raise NotImplementedError("HDF5 catalogue files are outside the hot-path scope (SURVEY.md section 8f)")
        if not isinstance(init_parms, dict):
            raise TypeError("init_parms must be a dictionary")
        p = init_parms
        self.location = NP.asarray(p["location"], dtype=NP.float64).reshape(-1, 2)
        nsrc = self.location.shape[0]
        self.name = p.get("name", NP.arange(nsrc).astype(str))
        self.frequency = NP.asarray(p.get("frequency", [150e6]), dtype=NP.float64).reshape(1, -1)
        self.coords = p.get("coords", "radec")
        self.epoch = p.get("epoch", "J2000")
        self.spec_type = p.get("spec_type", "func")
        self.src_shape = None
        if p.get("src_shape", None) is not None:
            self.src_shape = NP.asarray(p["src_shape"], dtype=NP.float64).reshape(nsrc, -1)
        self.spec_parms = {}
        self.spectrum = None
        if self.spec_type == "func":
            sp = p["spec_parms"]
            names = sp.get("name", NP.repeat("power-law", nsrc))
            if NP.any(NP.asarray(names) != "power-law"):
                raise NotImplementedError("only power-law functional spectra are on the hot path")

            def col(key, default):
                v = NP.asarray(sp.get(key, default), dtype=NP.float64).ravel()
                return NP.repeat(v, nsrc) if v.size == 1 else v

            self.spec_parms = {
                "name": NP.asarray(names),
                "power-law-index": col("power-law-index", 0.0),
                "freq-ref": col("freq-ref", self.frequency.ravel()[0]),
                "flux-scale": col("flux-scale", 1.0),
                "flux-offset": col("flux-offset", 0.0),
                "freq-width": col("freq-width", 0.0),
            }
            for key in ("power-law-index", "freq-ref", "flux-scale", "flux-offset"):
                if self.spec_parms[key].size != nsrc:
                    raise ValueError("spec_parms['{0}'] must have one entry per source".format(key))
        elif self.spec_type == "spectrum":
            self.spectrum = NP.asarray(p["spectrum"], dtype=NP.float64).reshape(nsrc, -1)
            if self.spectrum.shape[1] != self.frequency.size:
                raise ValueError("spectrum must be [nsrc, len(frequency)]")
        else:
            raise ValueError("spec_type must be 'func' or 'spectrum'")

    def generate_spectrum(self, ind=None, frequency=None, interp_method="pchip"):
        """Host evaluation of the catalogue spectra (convenience; ``observe`` evaluates power laws
        on the GPU inside ``pb200_amp_table``).  Tabulated spectra are interpolated here."""
        freq = self.frequency.ravel() if frequency is None else NP.asarray(frequency, dtype=NP.float64).ravel()
        sel = slice(None) if ind is None else NP.asarray(ind)
        if self.spec_type == "func":
            sp = self.spec_parms
            return (sp["flux-offset"][sel].reshape(-1, 1) + sp["flux-scale"][sel].reshape(-1, 1)
                    * (freq.reshape(1, -1) / sp["freq-ref"][sel].reshape(-1, 1)) ** sp["power-law-index"][sel].reshape(-1, 1))
        spec = self.spectrum[sel]
        f0 = self.frequency.ravel()
        if f0.size == freq.size and NP.allclose(f0, freq, rtol=0, atol=1e-6):
            return spec
        from scipy import interpolate
        if interp_method == "pchip" and f0.size > 1:
            return interpolate.PchipInterpolator(f0, spec, axis=1)(freq)
        return interpolate.interp1d(f0, spec, axis=1, kind="linear", bounds_error=False, fill_value="extrapolate")(freq)

    def subset(self, indices=None, axis="position"):
        if indices is None:
            return self
        indices = NP.asarray(indices)
        parms = {"location": self.location[indices], "coords": self.coords, "epoch": self.epoch,
                 "frequency": self.frequency, "spec_type": self.spec_type, "name": NP.asarray(self.name)[indices],
                 "src_shape": None if self.src_shape is None else self.src_shape[indices]}
        if self.spec_type == "func":
            parms["spec_parms"] = {k: v[indices] for k, v in self.spec_parms.items()}
        else:
            parms["spectrum"] = self.spectrum[indices]
        return SkyModel(init_parms=parms)
