"""Host-side mirror of the reference's ``prisim/interferometry.py`` for the hot path.

``InterferometerArray`` keeps the reference's constructor and method signatures
(interferometry.py:5140-5145, :5874-5878, :6414-6417, :6661, :6697, :8052) and the attribute names
a PRISim user reads afterwards, but every number is produced by the CUDA
kernels behind ``include/prisim_b200.h``.  Per-snapshot products live on the GPU as
[nbl, nchan] tensors; the reference-shaped [nbl, nchan, nsnap] numpy arrays are materialised on
attribute access.

Deliberate deviations from the reference (documented in DESIGN.md):
  * sky coordinates 'radec' are turned into hour angles with HA = LST - RA (the reference's own
    legacy path, interferometry.py:4482); astropy's apparent-place pipeline (:6174-6180) is an
    input producer outside the parity boundary.  'altaz' sky coordinates are accepted (the
    reference raises NameError, Appendix C #2).
  * ``geometric_delays`` is never materialised (Appendix C #4); ``memsave`` is ignored (fp64 path
    is the parity target); ``gradient_mode='baseline'`` fills ``gradient`` (three more phase sums).
  * noise uses a counter-based Philox generator instead of numpy's global state (Appendix C #16).
"""
from __future__ import annotations

import os
import warnings
from collections import OrderedDict

import numpy as NP
import scipy.constants as FCNST
import torch

from . import engine
from . import geometry as GEOM
from . import primary_beams as PB


class SimpleTime(object):
    """Duck-type of the astropy ``Time`` the reference's ``observe`` consumes: it only calls
    ``timeobj.sidereal_time('apparent').deg`` (:6113) and ``timeobj.jd`` (:6395)."""

    class _Angle(object):
        def __init__(self, deg):
            self.deg = deg

    def __init__(self, jd, lst_deg):
        self.jd = float(jd)
        self._lst = float(lst_deg)

    def sidereal_time(self, kind="apparent", longitude=None):
        return SimpleTime._Angle(self._lst)


################################################################################
# Layout helpers (host-side input producers; O(N_ant^2))
################################################################################

_BCAST_SHAPES = lambda nbl, nchan, ntimes: [(a, b, c) for a in (1, nbl) for b in (1, nchan) for c in (1, ntimes)]


def _noise_input(name, val, nbl, nchan, ntimes):
    if not isinstance(val, (int, float, list, NP.ndarray)):
        raise TypeError("Input {0} must be a scalar, float, list or numpy array".format(name))
    arr = NP.asarray(val, dtype=NP.float64).reshape(1, 1, 1) if isinstance(val, (int, float)) else NP.asarray(val, dtype=NP.float64)
    if NP.any(arr < 0.0):
        raise ValueError("Value(s) in {0} cannot be negative".format(name))
    if arr.shape not in _BCAST_SHAPES(nbl, nchan, ntimes):
        raise IndexError("{0} specified has incompatible dimensions".format(name))
    return arr


def fresh_noise_seed():
    """A 63-bit seed from the operating system: the default of ``InterferometerArray(noise_seed=None)`` and
    ``generateNoise(seed=None)``, so that separate objects / calls draw independent noise as the reference's
    ``NP.random.randn`` does (interferometry.py:6693, :329).  Pass an explicit seed for reproducible or sharded runs."""
    return int.from_bytes(os.urandom(8), "little") >> 1


def _realisation_seed(seed, realisation):
    """Philox key of the `realisation`-th noise draw of one object: the seed itself for the first draw, a splitmix64
    hash of (seed, realisation) afterwards -- repeated ``generate_noise()`` calls give fresh, reproducible noise."""
    seed = int(seed) & 0xFFFFFFFFFFFFFFFF
    if realisation == 0:
        return seed
    z = (seed + 0x9E3779B97F4A7C15 * int(realisation)) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return z ^ (z >> 31)


def thermalNoiseRMS(A_eff, df, dt, Tsys, nbl=1, nchan=1, ntimes=1, flux_unit="Jy", eff_Q=1.0):
    """Thermal-noise rms of a complex visibility, same call as interferometry.py:89-230: 2 k Tsys / (A_eff eff_Q sqrt(df dt))
    in Jy, or Tsys / (eff_Q sqrt(df dt)) in K; inputs are scalars or arrays broadcastable to (nbl, nchan, ntimes)."""
    if not isinstance(df, (int, float)):
        raise TypeError("Input channel resolution must be a scalar")
    if not isinstance(dt, (int, float)):
        raise TypeError("Input time resolution must be a scalar")
    for name, val in (("nbl", nbl), ("nchan", nchan), ("ntimes", ntimes)):
        if not isinstance(val, int):
            raise TypeError("Input {0} must be an integer".format(name))
        if val <= 0:
            raise ValueError("Input {0} must be positive".format(name))
    Tsys = _noise_input("Tsys", Tsys, nbl, nchan, ntimes)
    A_eff = _noise_input("A_eff", A_eff, nbl, nchan, ntimes)
    eff_Q = _noise_input("eff_Q", eff_Q, nbl, nchan, ntimes)
    if not isinstance(flux_unit, str):
        raise TypeError("Input flux_unit must be a string")
    if flux_unit.lower() not in ["k", "jy"]:
        raise ValueError("Input flux_unit must be set to K or Jy")
    if flux_unit.lower() == "k":
        return Tsys / eff_Q / NP.sqrt(float(dt) * float(df))
    return 2.0 * FCNST.k / NP.sqrt(float(dt) * float(df)) * (Tsys / A_eff / eff_Q) / 1.0e-26


def generateNoise(noiseRMS=None, A_eff=None, df=None, dt=None, Tsys=None, nbl=1, nchan=1, ntimes=1, flux_unit="Jy", eff_Q=1.0,
                  seed=None, device=None):
    """Complex thermal noise rms/sqrt(2) (N + iN) of shape (nbl, nchan, ntimes), same call as interferometry.py:236-329
    plus `seed` / `device`: the deviates come from the device Philox generator (``pb200_noise``), one launch per time.
    seed=None draws a fresh seed per call (independent realisations, like the reference's NP.random.randn)."""
    if seed is None:
        seed = fresh_noise_seed()
    if noiseRMS is None:
        noiseRMS = thermalNoiseRMS(A_eff, df, dt, Tsys, nbl=nbl, nchan=nchan, ntimes=ntimes, flux_unit=flux_unit, eff_Q=eff_Q)
    else:
        noiseRMS = _noise_input("noiseRMS", noiseRMS, nbl, nchan, ntimes)
    rms = NP.broadcast_to(noiseRMS, (nbl, nchan, ntimes))
    dev = engine._dev(device)
    one = engine._f64([1.0], dev)
    out = NP.empty((nbl, nchan, ntimes), dtype=NP.complex128)
    for t in range(ntimes):
        # the K form of the kernel, rms = Tsys / eff_Q / sqrt(df dt), with Tsys := rms and unit eff_Q, df, dt
        _, nz, _ = engine.noise(None, engine._f64(NP.ascontiguousarray(rms[:, :, t]), dev), None, one, 1.0, 1.0, int(seed), nbl, nchan,
                                snapshot=t, flux_unit_k=True, want=("noise",), device=dev)
        out[:, :, t] = nz.cpu().numpy()
    return out


def hexagon_generator(spacing, n_total=None, n_side=None, orientation=None, center=None):
    """Hexagonal antenna layout with the call and the antenna ORDER of interferometry.py:857-989: for every offset
    i = 1 .. n_side-1 the row above the centre line, then its mirror below, and the centre line last; centred on the
    mean, rotated by `orientation` degrees, scaled by `spacing`, shifted by `center`.  Returns (xy [n,2], labels)."""
    if n_total is None and n_side is None:
        raise NameError("n_total or n_side must be provided")
    if n_side is None:
        # n_total = 3 n^2 - 3 n + 1  =>  n = (3 + sqrt(12 n_total - 3)) / 6
        n_side = int(round((3.0 + NP.sqrt(12.0 * n_total - 3.0)) / 6.0))
        if n_side < 1 or 3 * n_side ** 2 - 3 * n_side + 1 != n_total:
            raise ValueError("n_total is not a valid number for a hexagonal array")
    else:
        if not isinstance(n_side, (int, NP.integer)):
            raise TypeError("n_side must be an integer")
        if n_side <= 0:
            raise ValueError("n_side must be positive")
        n_total = 3 * n_side ** 2 - 3 * n_side + 1
    width = 2 * n_side - 1
    dx, dy = NP.cos(NP.pi / 3), NP.sin(NP.pi / 3)          # the reference's own constants (0.5000000000000001, 0.866...)
    blocks = []
    for i in range(1, n_side):
        xs = NP.arange(width - i, dtype=NP.float64) + i * dx
        ys = NP.full(width - i, i * dy)
        blocks += [NP.column_stack((xs, ys)), NP.column_stack((xs, -ys))]
    blocks.append(NP.column_stack((NP.arange(width, dtype=NP.float64), NP.zeros(width))))
    xy = NP.concatenate(blocks, axis=0)
    xy = xy - NP.mean(xy, axis=0, keepdims=True)
    if orientation is not None:
        ang = NP.radians(orientation)
        xy = NP.dot(xy, NP.asarray([[NP.cos(ang), -NP.sin(ang)], [NP.sin(ang), NP.cos(ang)]]).T)
    xy = xy * spacing
    if center is not None:
        xy = xy + center
    return xy, [str(i) for i in range(n_total)]


def baseline_generator(antenna_locations, ant_label=None, ant_id=None, auto=False, conjugate=False):
    """All antenna pairs, mirrors interferometry.py:1184-1370 for numpy input: b = r_j - r_i for
    j > i with i as the outer loop (:1355-1358).  Returns (baselines [nbl,3], labels, ids)."""
    ant = NP.asarray(antenna_locations, dtype=NP.float64)
    if ant.ndim == 1:
        ant = ant.reshape(-1, 1)
    if ant.shape[1] < 3:
        ant = NP.hstack((ant, NP.zeros((ant.shape[0], 3 - ant.shape[1]))))
    n = ant.shape[0]
    if ant_label is None:
        ant_label = NP.asarray([str(i) for i in range(n)])                 # :1302
    ant_label = NP.asarray(ant_label)
    if ant_id is None:
        ant_id = NP.arange(n)
    ii, jj = NP.triu_indices(n, k=0 if auto else 1)          # i outer, j inner, j > i (or >=)
    if conjugate:
        i2, j2 = NP.tril_indices(n, k=-1)                    # j < i, i outer
        ii, jj = NP.concatenate((ii, i2)), NP.concatenate((jj, j2))
    bl = ant[jj] - ant[ii]
    maxlen = max(len(str(a)) for a in ant_label)
    labels = NP.asarray(list(zip(ant_label[jj], ant_label[ii])), dtype=[("A2", "U{0}".format(maxlen)), ("A1", "U{0}".format(maxlen))])
    ids = NP.asarray(list(zip(NP.asarray(ant_id)[jj], NP.asarray(ant_id)[ii])), dtype=[("A2", int), ("A1", int)])
    return bl, labels, ids


def orient_and_sort_baselines(bl, labels=None):
    """run through the reference's conjugation and ordering step (interferometry.py:1868-1883):
    flip baselines whose orientation falls outside (-67.5, 112.5] degrees, then stable-sort by
    length."""
    bl = NP.array(bl, dtype=NP.float64)
    blo = NP.angle(bl[:, 0] + 1j * bl[:, 1], deg=True)
    neg = (blo < -67.5) | (blo > 112.5)
    bl[neg] = -1.0 * bl[neg]
    if labels is not None:
        labels = NP.array(labels)
        lab2 = labels.copy()
        lab2["A2"][neg], lab2["A1"][neg] = labels["A1"][neg], labels["A2"][neg]
        labels = lab2
    order = NP.argsort(NP.sqrt(NP.sum(bl ** 2, axis=1)), kind="mergesort")
    return bl[order], (None if labels is None else labels[order]), order


def uniq_baselines(baseline_locations, redundant=None):
    """Same call and 4-tuple as interferometry.py:1373-1463: (selected baselines [nu,3], their first-occurrence indices,
    counts, list of the indices of all occurrences of each).  Two baselines are the same when length (0.01 m), zenith
    angle and orientation folded into [0, 180) degrees (0.001 arcsec) agree -- b and -b are one baseline -- and the
    result is ordered like the reference orders its formatted keys.  redundant=None: every unique baseline; True:
    only those occurring more than once; False: only those occurring once."""
    if not isinstance(baseline_locations, NP.ndarray):
        raise TypeError("baseline_locations must be a numpy array")
    if redundant is not None and not isinstance(redundant, bool):
        raise TypeError('keyword "redundant" must be set to None or a boolean value')
    b3 = NP.zeros((baseline_locations.shape[0], 3))
    ncol = min(3, baseline_locations.shape[1])
    b3[:, :ncol] = baseline_locations[:, :ncol]
    length = NP.linalg.norm(b3, axis=1)
    za_arcsec = 3.6e3 * NP.degrees(NP.arccos(b3[:, 2] / length))
    orient = NP.degrees(NP.arctan2(b3[:, 1], b3[:, 0]))
    orient = NP.where(orient >= 180.0, orient - 180.0, NP.where(orient < 0.0, orient + 180.0, orient))
    keys = NP.asarray(["%.2f_%.3f_%.3f" % t for t in zip(length, za_arcsec, 3.6e3 * orient)])
    _, first, inverse, count = NP.unique(keys, return_index=True, return_inverse=True, return_counts=True)
    pick = NP.arange(first.size) if redundant is None else NP.flatnonzero(count > 1 if redundant else count == 1)
    occurrences = [NP.flatnonzero(inverse == g).tolist() for g in pick]
    counts = count[pick] if redundant is not False else NP.ones(pick.size)
    return b3[first[pick], :], first[pick], counts, occurrences


def baseline_groups(labels, baseline_locations):
    """Unique baselines of a (redundant) array with the bookkeeping ``duplicate_measurements`` needs, in the dictionary
    formats getBaselineInfo builds (interferometry.py:1999-2007): returns (unique labels, unique baselines [nu,3],
    {'groups': {label tuple: labels of all baselines redundant with it}, 'reversemap': {label tuple: group key}})."""
    labels = NP.asarray(labels)
    ubl, first, _, occ = uniq_baselines(NP.asarray(baseline_locations, dtype=NP.float64))
    order = NP.argsort(first, kind="stable")                          # keep the input order of first occurrences
    groups, reverse = {}, {}
    for g in order:
        key = tuple(labels[first[g]].tolist()) if labels[first[g]].shape or labels.dtype.names else (labels[first[g]].item(),)
        members = labels[NP.asarray(occ[g])]
        groups[key] = members
        for lbl in members:
            reverse[tuple(lbl.tolist()) if (lbl.shape or labels.dtype.names) else (lbl.item(),)] = NP.asarray([labels[first[g]]], dtype=labels.dtype)
    return labels[first[order]], ubl[order], {"groups": groups, "reversemap": reverse}


def getBaselineGroupKeys(inp_labels, blgroups_reversemap):
    """Same call as interferometry.py:2017-2095: for every label tuple the key of the redundant group it belongs to
    (looked up as given, else reversed) and whether it had to be reversed; (None, None) when it is in neither form."""
    if not isinstance(blgroups_reversemap, dict):
        raise TypeError("Input blgroups_reversemap must be a dictionary")
    if not isinstance(inp_labels, list):
        inp_labels = [inp_labels]
    keys, flipped = [], []
    for lbl in inp_labels:
        hit = None
        for cand, flip in ((lbl, False), (lbl[::-1], True)):
            if cand in blgroups_reversemap:
                hit = (blgroups_reversemap[cand], flip)
                break
        if hit is None:
            keys.append(None); flipped.append(None)
            continue
        val = hit[0]
        if isinstance(val, NP.ndarray):
            keys.append(tuple(val[0].tolist()) if val.dtype.names else tuple(val[0]))
        elif isinstance(val, tuple):
            keys.append(val)
        else:
            raise TypeError("Invalid type found in blgroups_reversemap")
        flipped.append(hit[1])
    return keys, flipped


def getBaselinesInGroups(inp_labels, blgroups_reversemap, blgroups):
    """Same call as interferometry.py:2100-2165: the label arrays of the redundant groups containing the given labels."""
    if not isinstance(blgroups, dict):
        raise TypeError("Input blgroups must be a dictionary")
    keys, flipped = getBaselineGroupKeys(inp_labels, blgroups_reversemap)
    return [None if k is None else blgroups[k] for k in keys], flipped


################################################################################

def _validate_2d(name, val, nbl, nchan):
    val = NP.asarray(val, dtype=NP.float64)
    if val.size == nbl:
        return NP.repeat(val.reshape(-1, 1), nchan, axis=1)
    if val.size == nchan:
        return NP.repeat(val.reshape(1, -1), nbl, axis=0)
    if val.size == nbl * nchan:
        return val.reshape(-1, nchan)
    raise ValueError("{0} incompatible with the number of interferometers and/or frequency channels.".format(name))


class ROI_parameters(object):
    """Drop-in for ``prisim.interferometry.ROI_parameters`` (interferometry.py:3905-4617) in memory: per snapshot the
    catalogue indices inside the region of interest and the primary-beam table [n_roi, nchan] that
    ``InterferometerArray.observe(roi_info={'ind', 'pbeam'})`` consumes.  The beam is evaluated on the GPU
    (``primary_beam_generator`` -> ``pb200_amp_table``).  ``save`` / ``init_file`` exchange the tables through the same
    FITS layout as the reference (:4621-4723 / :4080-4205) using the plain-numpy writer in ``fits_min`` (astropy is not
    available here); on the GPU path the tables normally stay in memory and this file is only an interchange format."""

    def __init__(self, init_file=None, device=None):
        self.device = device
        if init_file is not None:
            self._load(init_file)
            return
        self.skymodel = None
        self.freq = None
        self.freq_scale = None
        self.telescope = None
        self.info = {"radius": [], "center": [], "ind": [], "pbeam": [], "center_coords": None}     # :4170-4176
        self.pinfo = []
        self.device = device

    def _load(self, init_file):
        """interferometry.py:4080-4205: rebuild the object from a file written by ``save`` (or by the reference)."""
        from . import fits_min as F
        try:
            hdus = F.read(init_file)
        except IOError:
            raise IOError("init_file provided but could not open the initialization file.")
        h0 = hdus[0][0]
        ext = OrderedDict((str(F.header_get(h, "EXTNAME", "")).upper(), (h, d)) for h, d in hdus[1:])
        n_obs = int(F.header_get(h0, "n_obs"))
        self.skymodel, self.freq_scale = None, "Hz"
        self.info = {"radius": [], "center": [], "ind": [], "pbeam": [], "center_coords": None}
        tel = {}
        if F.header_get(h0, "telescope") is not None:
            tel["id"] = F.header_get(h0, "telescope")
        tel["latitude"] = F.header_get(h0, "latitude", None)
        tel["longitude"] = F.header_get(h0, "longitude", 0.0)
        tel["altitude"] = F.header_get(h0, "altitude", 0.0)
        for key, name in (("shape", "element_shape"), ("size", "element_size"), ("ocoords", "element_ocoords")):
            if F.header_get(h0, name) is None:
                raise KeyError("Antenna {0} not found in the init_file header".format(name.replace("_", " ")))
            tel[key] = F.header_get(h0, name)
        if "ANTENNA ELEMENT ORIENTATION" not in ext:
            raise KeyError('Extension named "orientation" not found in init_file.')
        tel["orientation"] = ext["ANTENNA ELEMENT ORIENTATION"][1].reshape(1, -1)
        if "ANTENNA ELEMENT LOCATIONS" in ext:
            tel["element_locs"] = ext["ANTENNA ELEMENT LOCATIONS"][1]
        tel["groundplane"] = F.header_get(h0, "ground_plane", None)
        for key in ("scale", "max"):
            v = F.header_get(h0, "ground_modify_" + key)
            if tel["groundplane"] is not None and v is not None:
                tel.setdefault("ground_modify", {})[key] = v
        self.telescope = tel
        if "FREQ" not in ext:
            raise KeyError('Extension named "FREQ" not found in init_file.')
        self.freq = ext["FREQ"][1]
        empty = NP.asarray([])
        self.info["ind"] = [ext["IND_{0:0d}".format(i)][1] if "IND_{0:0d}".format(i) in ext else empty for i in range(n_obs)]
        self.info["pbeam"] = [ext["PB_{0:0d}".format(i)][1] if "PB_{0:0d}".format(i) in ext else empty for i in range(n_obs)]
        self.pinfo = []
        if any(k.startswith("DELAYS_") or k.startswith("POINTING_CENTER_") for k in ext):
            for i in range(n_obs):
                p = {}
                if "DELAYS_{0:0d}".format(i) in ext:
                    h, d = ext["DELAYS_{0:0d}".format(i)]
                    p["delays"] = d
                    err = F.header_get(h, "delayerr")
                    if err is not None:
                        p["delayerr"] = None if err <= 0.0 else err
                if "POINTING_CENTER_{0:0d}".format(i) in ext:
                    h, d = ext["POINTING_CENTER_{0:0d}".format(i)]
                    p["pointing_center"] = d
                    if F.header_get(h, "pointing_coords") is None:
                        raise KeyError('Header of extension POINTING_CENTER_{0:0d} not found to contain key "pointing_coords" in init_file'.format(i))
                    p["pointing_coords"] = F.header_get(h, "pointing_coords")
                self.pinfo.append(p)

    def save(self, infile, tabtype="BinTableHDU", overwrite=False, verbose=True):
        """Same call and file layout as interferometry.py:4621-4723: ``infile + '.fits'`` with the telescope keywords in
        the primary header and one image extension per array (element orientation / locations, FREQ, IND_j, PB_j, DELAYS_j,
        POINTING_CENTER_j).  ``fits_min.getdata(file, 'IND_3')`` reads a table back like ``fits.getdata`` in run_prisim."""
        from . import fits_min as F
        if not isinstance(infile, str):
            raise TypeError("Output filename must be a string")
        tel = self.telescope
        hdr = OrderedDict()
        hdr["n_obs"] = (len(self.info["ind"]), "Number of observations")
        if "id" in tel:
            hdr["telescope"] = (tel["id"], "Telescope Name")
        hdr["element_shape"] = (tel["shape"], "Antenna element shape")
        hdr["element_size"] = (float(tel["size"]), "Antenna element size [m]")
        hdr["element_ocoords"] = (tel["ocoords"], "Antenna element orientation coordinates")
        if tel.get("latitude", None) is not None:
            hdr["latitude"] = (float(tel["latitude"]), "Latitude (in degrees)")
        hdr["longitude"] = (float(tel.get("longitude", 0.0) or 0.0), "Longitude (in degrees)")
        if tel.get("altitude", None) is not None:
            hdr["altitude"] = (float(tel["altitude"]), "Altitude (in m)")
        if tel.get("groundplane", None) is not None:
            hdr["ground_plane"] = (float(tel["groundplane"]), "Antenna element height above ground plane [m]")
            for key, text in (("scale", "Ground plane modification scale factor"), ("max", "Maximum ground plane modification")):
                if key in tel.get("ground_modify", {}):
                    hdr["ground_modify_" + key] = (float(tel["ground_modify"][key]), text)
        ext = [("ANTENNA ELEMENT ORIENTATION", NP.asarray(tel["orientation"], dtype=NP.float64), None)]
        if "element_locs" in tel:
            ext.append(("ANTENNA ELEMENT LOCATIONS", NP.asarray(tel["element_locs"], dtype=NP.float64), None))
        ext.append(("FREQ", NP.asarray(self.freq, dtype=NP.float64), None))
        for i in range(len(self.info["ind"])):
            if NP.asarray(self.info["ind"][i]).size > 0:
                ext.append(("IND_{0:0d}".format(i), NP.asarray(self.info["ind"][i]), None))
                ext.append(("PB_{0:0d}".format(i), NP.asarray(self.info["pbeam"][i]), None))
            if self.pinfo and i < len(self.pinfo) and self.pinfo[i] is not None:
                p = self.pinfo[i]
                if "delays" in p:
                    err = p.get("delayerr", None)
                    ext.append(("DELAYS_{0:0d}".format(i), NP.asarray(p["delays"], dtype=NP.float64),
                                {"delayerr": (0.0 if err is None else float(err), "Jitter in delays [s]")}))
                if "pointing_center" in p:
                    if "pointing_coords" not in p:
                        raise KeyError('Key "pointing_coords" not found in attribute pinfo.')
                    ext.append(("POINTING_CENTER_{0:0d}".format(i), NP.asarray(p["pointing_center"], dtype=NP.float64),
                                {"pointing_coords": (p["pointing_coords"], "Pointing coordinate system")}))
        F.write(infile + ".fits", hdr, ext, overwrite=overwrite)
        if verbose:
            print("\tRegions of interest information written successfully to FITS file on disk:\n\t\t{0}\n".format(infile + ".fits"))

    def _altaz(self, skymodel, lst):
        coords = getattr(skymodel, "coords", None)
        loc = NP.asarray(skymodel.location, dtype=NP.float64)
        if coords in ("radec", "hadec") and self.telescope.get("latitude", None) is None:
            raise ValueError("Latitude of the observatory must be provided.")
        if coords == "radec":          # HA = LST - RA (same deviation from the astropy route as observe())
            if lst is None:
                raise ValueError("LST must be provided.")
            return GEOM.hadec2altaz(NP.stack(((lst - loc[:, 0]), loc[:, 1]), axis=1), self.telescope["latitude"], units="degrees")
        if coords == "hadec":
            return GEOM.hadec2altaz(loc, self.telescope["latitude"], units="degrees")
        if coords == "dircos":
            return GEOM.dircos2altaz(loc, units="degrees")
        if coords == "altaz":
            return loc
        raise KeyError("skycoords invalid or unspecified in skymodel")

    def append_settings(self, skymodel, freq, pinfo=None, lst=None, time_jd=None, roi_info=None, telescope=None, freq_scale="GHz"):
        """Same call as interferometry.py:4221-4224."""
        from .skymodel import SkyModel
        if self.freq is None:                                                          # :4396-4414
            if freq is None:
                raise ValueError("freq must be specified using a numpy array")
            if not isinstance(freq, NP.ndarray):
                raise TypeError("freq must be specified using a numpy array")
            scale = {None: 1.0, "hz": 1.0, "ghz": 1.0e9, "mhz": 1.0e6, "khz": 1.0e3}
            key = freq_scale.lower() if isinstance(freq_scale, str) else freq_scale
            if key not in scale:
                raise ValueError('Frequency units must be "GHz", "MHz", "kHz" or "Hz". If not set, it defaults to "Hz"')
            self.freq = NP.asarray(freq, dtype=NP.float64).ravel() * scale[key]
            self.freq_scale = "Hz"
        if self.telescope is None:                                                     # :4416-4420
            if not isinstance(telescope, dict):
                raise TypeError("Input telescope must be a dictionary.")
            self.telescope = telescope
        if skymodel is None:                                                           # :4422-4425
            self.info["pbeam"] += [NP.asarray([])]
            self.info["ind"] += [NP.asarray([])]
            self.pinfo += [None]
            return
        if not isinstance(skymodel, SkyModel):
            raise TypeError("skymodel should be an instance of class SkyModel.")
        self.skymodel = skymodel
        if roi_info is None:
            raise ValueError("roi_info dictionary must be set.")
        pbeam_input = False
        skypos_altaz = None
        if roi_info.get("ind", None) is not None:                                      # :4451-4495
            ind = NP.asarray(roi_info["ind"])
            self.info["ind"] += [ind]
            if ind.size > 0:
                if roi_info.get("pbeam", None) is not None:
                    try:
                        pb = NP.asarray(roi_info["pbeam"]).reshape(-1, self.freq.size)
                    except ValueError:
                        raise ValueError('Number of columns of primary beam in key "pbeam" of dictionary roi_info must be equal to number of frequency channels.')
                    if ind.size != pb.shape[0]:
                        raise ValueError('Number of elements in values in key "ind" and number of rows of values in key "pbeam" must be identical.')
                    self.info["pbeam"] += [NP.asarray(roi_info["pbeam"]).astype(NP.float32)]
                    pbeam_input = True
                if not pbeam_input:
                    skypos_altaz = self._altaz(skymodel, lst)
            if "radius" in roi_info:
                self.info["radius"] += [roi_info["radius"]]
            if "center" in roi_info:
                self.info["center"] += [roi_info["center"]]
        else:                                                                          # :4496-4552
            radius = 90.0 if roi_info.get("radius", None) is None else max(0.0, min(roi_info["radius"], 90.0))
            self.info["radius"] += [radius]
            lat = self.telescope.get("latitude", None)
            if roi_info.get("center", None) is None:
                self.info["center"] += [NP.asarray([90.0, 270.0]).reshape(1, -1)]
            else:
                c = NP.asarray(roi_info["center"], dtype=NP.float64).reshape(1, -1)
                cc = roi_info.get("center_coords", None)
                if cc == "dircos":
                    self.info["center"] += [GEOM.dircos2altaz(c, units="degrees")]
                elif cc == "altaz":
                    self.info["center"] += [c]
                elif cc == "hadec":
                    self.info["center"] += [GEOM.hadec2altaz(c, lat, units="degrees")]
                elif cc == "radec":
                    if lst is None:
                        raise KeyError("LST not provided for coordinate conversion")
                    self.info["center"] += [GEOM.hadec2altaz(NP.asarray([lst - c[0, 0], c[0, 1]]).reshape(1, -1), lat, units="degrees")]
                else:
                    raise ValueError("Invalid coordinate system specified for center")
            skypos_altaz = self._altaz(skymodel, lst)
            centre = self.info["center"][-1]
            if _sphdist(centre[0, 1], centre[0, 0], 270.0, 90.0) > 1e-2:             # :4546-4549 ROI centre is not the zenith
                # the reference hands (alt, az) to spherematch in its (lon, lat) slots; kept as written for parity
                ind = NP.where(_sphdist(centre[0, 0], centre[0, 1], skypos_altaz[:, 0], skypos_altaz[:, 1]) <= radius)[0]
            else:
                ind = NP.where(skypos_altaz[:, 0] >= 90.0 - radius)[0]                 # :4551
            self.info["ind"] += [ind]
        if self.info["center_coords"] is None and roi_info.get("center_coords", None) in ("altaz", "dircos", "hadec", "radec"):
            self.info["center_coords"] = roi_info["center_coords"]                     # :4554-4557
        if pbeam_input:
            return
        if pinfo is None:                                                              # :4559-4576
            raise ValueError("Pointing info dictionary pinfo must be specified.")
        pinfo = dict(pinfo)
        self.pinfo += [pinfo]
        pcoords = pinfo.get("pointing_coords", None)
        if pcoords is not None and pcoords not in ("dircos", "altaz"):
            if self.telescope.get("latitude", None) is None:
                raise ValueError("Latitude of the observatory must be provided.")
            pc = NP.asarray(pinfo["pointing_center"], dtype=NP.float64).reshape(1, -1)
            if pcoords == "radec":
                if lst is None:
                    raise ValueError("LST must be provided.")
                pc = NP.asarray([lst - pc[0, 0], pc[0, 1]]).reshape(1, -1)
            elif pcoords != "hadec":
                raise ValueError('pointing_coords in dictionary pinfo must be "dircos", "altaz", "hadec" or "radec".')
            pinfo["pointing_center"] = GEOM.hadec2altaz(pc, self.telescope["latitude"], units="degrees")
            pinfo["pointing_coords"] = "altaz"
        ind = self.info["ind"][-1]
        if ind.size == 0:
            self.info["pbeam"] += [NP.asarray([])]
            return
        reffreq = roi_info.get("pbeam_reffreq", self.freq[self.freq.size // 2])        # :4578-4588
        fcomp = self.freq if roi_info.get("pbeam_chromaticity", False) else NP.asarray(reffreq, dtype=NP.float64).reshape(-1)
        if self.telescope.get("id", None) == "mwa_tools":
            raise NotImplementedError("the external MWA_Tools beam is not on the hot path")
        pbeam = PB.primary_beam_generator(skypos_altaz[ind, :], fcomp, self.telescope, freq_scale="Hz", skyunits="altaz",
                                          pointing_info=pinfo, device=self.device)
        self.info["pbeam"] += [pbeam.astype(NP.float64) * NP.ones(self.freq.size).reshape(1, -1)]   # :4615


def _sphdist(lon1, lat1, lon2, lat2):
    """Great-circle separation in degrees (``GEOM.sphdist`` of astroutils), haversine form."""
    lon1, lat1, lon2, lat2 = [NP.radians(NP.asarray(v, dtype=NP.float64)) for v in (lon1, lat1, lon2, lat2)]
    a = NP.sin(0.5 * (lat2 - lat1)) ** 2 + NP.cos(lat1) * NP.cos(lat2) * NP.sin(0.5 * (lon2 - lon1)) ** 2
    with NP.errstate(invalid="ignore"):        # "latitudes" beyond 90 deg (the alt/az swap above) can make a < 0: NaN, never matched
        return NP.degrees(2.0 * NP.arcsin(NP.minimum(1.0, NP.sqrt(a))))


class InterferometerArray(object):
    """Drop-in for ``prisim.interferometry.InterferometerArray`` on the visibility hot path."""

    def __init__(self, labels, baselines, channels, telescope=None, eff_Q=0.89, latitude=34.0790, longitude=0.0,
                 altitude=0.0, skycoords="radec", A_eff=NP.pi * (25.0 / 2) ** 2, pointing_coords="hadec", layout=None,
                 blgroupinfo=None, baseline_coords="localenu", freq_scale=None, gaininfo=None, init_file=None,
                 simparms_file=None, device=None, bl_offset=0, nbl_total=None, noise_seed=None, bl_step=1):
        if init_file is not None:
            raise NotImplementedError("loading saved simulations is outside the hot-path scope (SURVEY.md section 8f)")
        if gaininfo is not None:
            raise NotImplementedError("gain tables are outside the hot-path scope; unity gains are used (interferometry.py:6707)")
        self.baselines = NP.asarray(baselines, dtype=NP.float64)                       # :5668-5685
        if self.baselines.ndim == 1:
            if self.baselines.size == 2:
                self.baselines = NP.hstack((self.baselines.reshape(1, -1), NP.zeros((1, 1))))
            elif self.baselines.size == 3:
                self.baselines = self.baselines.reshape(1, -1)
            else:
                raise ValueError("Baseline(s) must be a 2- or 3-column array.")
        elif self.baselines.ndim == 2:
            if self.baselines.shape[1] == 2:
                self.baselines = NP.hstack((self.baselines, NP.zeros(self.baselines.shape[0]).reshape(-1, 1)))
            elif self.baselines.shape[1] != 3:
                raise ValueError("Baseline(s) must be a 2- or 3-column array")
        else:
            raise ValueError("Baseline(s) array contains more than 2 dimensions.")
        self.baseline_lengths = NP.sqrt(NP.sum(self.baselines ** 2, axis=1))
        self.baseline_orientations = NP.angle(self.baselines[:, 0] + 1j * self.baselines[:, 1])
        self.projected_baselines = None
        if not isinstance(labels, (list, tuple, NP.ndarray)):                          # :5690-5695
            raise TypeError("Interferometer array labels must be a list or tuple of unique identifiers")
        if len(labels) != self.baselines.shape[0]:
            raise ValueError("Number of labels do not match the number of baselines specified.")
        self.labels = labels
        self.simparms_file = simparms_file if isinstance(simparms_file, str) else None
        if isinstance(telescope, dict):                                                # :5703-5712
            self.telescope = telescope
        else:
            self.telescope = {"id": "vla", "shape": "dish", "size": 25.0, "ocoords": "altaz",
                              "orientation": NP.asarray([90.0, 270.0]).reshape(1, -1), "groundplane": None}
        self.layout = dict(layout) if isinstance(layout, dict) else {}
        self.blgroups = None
        self.bl_reversemap = None
        if blgroupinfo is not None:
            if not isinstance(blgroupinfo, dict):
                raise TypeError("Input blgroupinfo must be a dictionary")
            self.blgroups = blgroupinfo["groups"]
            self.bl_reversemap = blgroupinfo["reversemap"]
        self.latitude, self.longitude, self.altitude = latitude, longitude, altitude
        self.gradient_mode = None
        self._gradient = []                              # per snapshot [3, nbl, nchan] complex128 (gradient_mode='baseline')
        self.gaininfo = None
        scale = {None: 1.0, "hz": 1.0, "ghz": 1.0e9, "mhz": 1.0e6, "khz": 1.0e3}                # :5772-5786
        key = freq_scale.lower() if isinstance(freq_scale, str) else freq_scale
        if key not in scale:
            raise ValueError('Frequency units must be "GHz", "MHz", "kHz" or "Hz". If not set, it defaults to "Hz"')
        self.channels = NP.asarray(channels, dtype=NP.float64).ravel() * scale[key]
        nbl, nchan = self.baselines.shape[0], self.channels.size
        self.Tsysinfo = []
        self.flux_unit = "JY"
        self.timestamp = []
        self.t_acc = []
        self.t_obs = 0.0
        self.n_acc = 0
        self.pointing_center = NP.empty([1, 2])
        self.phase_center = NP.empty([1, 2])
        self.lst = []
        if isinstance(eff_Q, (int, float)):                                            # :5803-5822
            self.eff_Q = eff_Q * NP.ones((nbl, nchan))
        elif isinstance(eff_Q, (list, tuple, NP.ndarray)):
            eff_Q = NP.asarray(eff_Q)
            if NP.any(eff_Q < 0.0) or NP.any(eff_Q > 1.0):
                raise ValueError("One or more values of eff_Q found to be outside the range [0,1].")
            self.eff_Q = _validate_2d("Efficiency values of interferometers", eff_Q, nbl, nchan)
        else:
            raise TypeError("Efficiency values of interferometers must be provided as a scalar, list, tuple or numpy array.")
        if isinstance(A_eff, (int, float)):                                            # :5824-5843
            if A_eff < 0.0:
                raise ValueError("Negative value for effective area is invalid.")
            self.A_eff = A_eff * NP.ones((nbl, nchan))
        elif isinstance(A_eff, (list, tuple, NP.ndarray)):
            A_eff = NP.asarray(A_eff)
            if NP.any(A_eff < 0.0):
                raise ValueError("One or more values of A_eff found to be negative.")
            self.A_eff = _validate_2d("Effective area(s) of interferometers", A_eff, nbl, nchan)
        else:
            raise TypeError("Effective area(s) of interferometers must be provided as a scalar, list, tuple or numpy array.")
        self.freq_resolution = self.channels[1] - self.channels[0] if nchan > 1 else 1.0       # :5847
        self.lags = None
        self.obs_catalog_indices = []
        self.geometric_delays = []          # never materialised (Appendix C #4)
        if pointing_coords not in ("radec", "hadec", "altaz"):                         # :5855-5859
            raise ValueError('Pointing center of the interferometer must be "radec", "hadec" or "altaz". Check inputs.')
        self.pointing_coords = pointing_coords
        self.phase_center_coords = pointing_coords
        if skycoords not in ("radec", "hadec", "altaz"):                               # :5861-5864
            raise ValueError('Sky coordinates must be "radec", "hadec" or "altaz". Check inputs.')
        self.skycoords = skycoords
        if baseline_coords not in ("equatorial", "localenu"):                          # :5866-5869
            raise ValueError('Baseline coordinates must be "equatorial" or "local". Check inputs.')
        self.baseline_coords = baseline_coords

        # ---- device state ----
        self.device = engine._dev(device)
        self.bl_offset = int(bl_offset)                  # position of this shard in the full baseline list: local row b is global
        self.bl_step = int(bl_step)                      # baseline bl_offset + b * bl_step (noise is keyed by the global index)
        self.nbl_total = nbl if nbl_total is None else int(nbl_total)
        # None: a fresh seed per object (recorded here), so that separate arrays (sub-bands, polarisations, Monte-Carlo
        # repeats) draw independent noise like the reference's global numpy generator; pass a seed for reproducible or
        # baseline-sharded runs (sharding.make_sharded_array shares rank 0's).  Every generate_noise() call advances
        # a realisation counter folded into the Philox key.
        self.noise_seed = fresh_noise_seed() if noise_seed is None else int(noise_seed)
        self._noise_realisation = 0
        # 'fp32': fp32 phasors/amplitudes everywhere (fastest); 'fp64': fp64 kernel + fp64 amplitude table;
        # 'auto': fp32 first; every baseline whose visibilities are a strongly cancelling sum
        # (rms_b < cancel_ratio * incoherent norm) is recomputed by the fp64 kernel, and a sample of the
        # remaining baselines is audited against fp64 -- if the audit misses the tolerance the whole
        # snapshot (and the following ones) runs in fp64 (DESIGN.md K1 "precision control")
        self.precision = "auto"
        self.cancel_ratio = 0.3                          # measured on config 2 (tools/err_c2.py): max |dV| = 2.1e-6 A2 over all 6e7 cells -> < 0.8e-5 rms_b above 0.27
        self.skyvis_method = "auto"                      # fp32 kernel variant (engine.skyvis method; A/B measurements)
        self.sort_by_brightness = True                   # feed the phase sum the brightest sources first (fp32 rounding)
        self.audit_baselines = 32                        # un-flagged baselines re-done in fp64 and compared per snapshot
        self.audit_tolerance = 0.8e-5                    # max |dV|/rms_b on the audited baselines before fp64 takes over
        self.precision_report = []                       # per snapshot: baselines recomputed in fp64
        self._fp64_sticky = False
        self.cache_sky = True                            # keep catalogue arrays resident between snapshots
        self._sky_cache = {}
        self._d_bl = None
        self._skyvis = []                                # per snapshot [nbl,nchan] complex128 CUDA tensors
        self._vis = []
        self._noise = []
        self._rms = []
        self._bp = []                                    # per snapshot [nchan] or [nbl,nchan] float64 CUDA
        self._bp_wts = None                              # callable t -> tensor, set by delay_transform
        self._Tsys = []
        self._lag = {}                                   # product name -> list of [nbl,nout] tensors
        self._d_aeff = None
        self._d_effq = None
        # [nbl, nchan] complex128 CUDA tensor the NEXT observe() writes its visibilities into (consumed by that call).
        # sharding.ShardedObserver points it at this rank's rows of the writing rank's buffer (NVLink peer memory).
        self.next_skyvis_out = None

    # ------------------------------------------------------------------ helpers
    def _dev_str(self):
        return "cuda:{0}".format(self.device)

    def _compact(self, arr2d):
        """[nbl,nchan] host array -> device tensor, collapsed to [nchan] when all rows agree."""
        arr2d = NP.asarray(arr2d, dtype=NP.float64)
        if arr2d.shape[0] > 1 and NP.all(arr2d == arr2d[0:1]):
            return engine._f64(arr2d[0], self.device)
        return engine._f64(arr2d, self.device)

    def _stack(self, lst, expand=False):
        """list of per-snapshot device tensors -> numpy [nbl, n, nsnap] (reference layout)."""
        if not lst:
            return None
        nbl = self.baselines.shape[0]
        out = []
        for t in lst:
            if expand and t.ndim == 1:
                t = t.unsqueeze(0).expand(nbl, t.shape[0])
            out.append(t.cpu().numpy())
        return NP.stack(out, axis=2)

    # reference-shaped views (host numpy, materialised on access)
    @property
    def skyvis_freq(self):
        return self._stack(self._skyvis)

    @property
    def vis_freq(self):
        return self._stack(self._vis)

    @property
    def vis_noise_freq(self):
        return self._stack(self._noise)

    @property
    def vis_rms_freq(self):
        return self._stack(self._rms)

    @property
    def bp(self):
        if not self._bp:
            return NP.ones((self.baselines.shape[0], self.channels.size))
        return self._stack(self._bp, expand=True)

    @property
    def bp_wts(self):
        if not self._bp:
            return NP.ones((self.baselines.shape[0], self.channels.size))
        if self._bp_wts is None:
            return NP.ones((self.baselines.shape[0], self.channels.size, len(self._bp)))
        base = getattr(self, "_drained", 0)
        return self._stack([self._bp_wts(base + t) for t in range(len(self._bp))], expand=True)

    @property
    def Tsys(self):
        if not self._Tsys:
            return NP.zeros((self.baselines.shape[0], self.channels.size))
        return self._stack(self._Tsys, expand=True)

    @property
    def skyvis_lag(self):
        return self._stack(self._lag.get("skyvis", []))

    @property
    def vis_lag(self):
        return self._stack(self._lag.get("vis", []))

    @property
    def vis_noise_lag(self):
        return self._stack(self._lag.get("noise", []))

    @property
    def lag_kernel(self):
        return self._stack(self._lag.get("kernel", []), expand=True)

    @property
    def gradient(self):
        """{} or {'baseline': [3, nbl, nchan, nsnap] complex128} (interferometry.py:4834-4845, :6386-6394)."""
        if self.gradient_mode is None or not self._gradient:
            return {}
        return {self.gradient_mode: torch.stack(self._gradient, dim=3).cpu().numpy()}

    def apply_gradients(self, gradient_mode=None, perturbations=None):
        """First-order perturbed visibilities from the stored gradient, same call as interferometry.py:6726-6819:
        perturbations = {'baseline': [..., 3, nbl] metres} -> [..., nbl, nchan, nsnap] complex128."""
        if gradient_mode is None:
            gradient_mode = self.gradient_mode
        if perturbations is None:
            perturbations = {gradient_mode: NP.zeros((1, 1, 1))}
        if self.gradient_mode is None or not self._gradient:
            raise AttributeError("No gradient attribute found")
        if not isinstance(perturbations, dict):
            raise TypeError("Input perturbations must be a dictionary")
        if not isinstance(gradient_mode, str):
            raise TypeError("Input gradient_mode must be a string")
        if gradient_mode not in ["baseline"]:
            raise KeyError("Specified gradient mode {0} not currently supported".format(gradient_mode))
        if gradient_mode not in perturbations:
            raise KeyError("{0} key not found in input perturbations".format(gradient_mode))
        if gradient_mode != self.gradient_mode:
            raise ValueError("Specified gradient mode {0} not found in attribute".format(gradient_mode))
        pert = perturbations[gradient_mode]
        if not isinstance(pert, NP.ndarray):
            raise TypeError("Perturbations must be specified as a numpy array")
        if pert.ndim == 2:
            pert = pert[NP.newaxis, ...]
        if pert.ndim < 2:
            raise ValueError("Perturbations must be two--dimensions or higher")
        inpshape = pert.shape
        pert = pert.reshape(-1, inpshape[-2], inpshape[-1])
        nbl, nchan = self.baselines.shape[0], self.channels.size
        if pert.shape[2] != nbl:
            raise ValueError("Number of {0} perturbations not equal to that in the gradient attribute".format(gradient_mode))
        if pert.shape[1] < 3:                                                          # :6801-6806
            warnings.warn("Only {0}-dimensional coordinates specified. Proceeding with zero perturbations in other coordinate axes.".format(pert.shape[1]))
            pert = NP.concatenate((pert, NP.zeros((pert.shape[0], 3 - pert.shape[1], nbl))), axis=1)
        elif pert.shape[1] > 3:                                                        # :6807-6809
            warnings.warn("{0}-dimensional coordinates specified. Proceeding with only the first three dimensions of coordinate axes.".format(3))
            pert = pert[:, :3, :]
        p = engine._f64(NP.ascontiguousarray(pert), self.device).to(torch.complex128)   # [nseed, 3, nbl]
        k = torch.as_tensor(-2.0j * NP.pi * self.channels / FCNST.c, device=self._dev_str())   # -i 2 pi / lambda   :6811
        out = torch.empty((pert.shape[0], nbl, nchan, len(self._gradient)), dtype=torch.complex128, device=self._dev_str())
        for t, G in enumerate(self._gradient):                                         # :6813
            out[:, :, :, t] = torch.einsum("sib,ibf->sbf", p, G) * k[None, None, :]
        return out.cpu().numpy().reshape(tuple(inpshape[:-2]) + (nbl, nchan, len(self._gradient)))

    def skyvis_freq_device(self, snapshot=-1):
        """The [nbl,nchan] complex128 CUDA tensor of one snapshot (no copy)."""
        return self._skyvis[snapshot]

    # ------------------------------------------------------------------ observe
    def _sky_to_device(self, skymodel):
        # identity cache; holding the object itself keeps its id() from being recycled
        if self.cache_sky and self._sky_cache.get("obj", None) is skymodel:
            return self._sky_cache["dev"]
        d = {"location": engine._f64(skymodel.location, self.device)}
        if getattr(skymodel, "spec_type", "func") == "func":
            sp = skymodel.spec_parms
            d["spec"] = {"flux_scale": engine._f64(sp["flux-scale"], self.device),
                         "index": engine._f64(sp["power-law-index"], self.device),
                         "freq_ref": engine._f64(sp["freq-ref"], self.device)}
            if "flux-offset" in sp and NP.any(NP.asarray(sp["flux-offset"]) != 0.0):
                d["spec"]["flux_offset"] = engine._f64(sp["flux-offset"], self.device)
        else:
            spectrum = skymodel.generate_spectrum(frequency=self.channels, interp_method="pchip")
            d["spec"] = {"spectrum": engine._f64(spectrum, self.device)}
        if getattr(skymodel, "src_shape", None) is not None:
            shp = NP.asarray(skymodel.src_shape, dtype=NP.float64)
            d["fwhm"] = engine._f64(NP.sqrt(shp[:, 0] * shp[:, 1]), self.device)            # :6267
        if self.cache_sky:
            self._sky_cache = {"obj": skymodel, "dev": d}
        return d

    def observe(self, timeobj, Tsysinfo, bandpass, pointing_center, skymodel, t_acc, pb_info=None,
                brightness_units=None, bpcorrect=None, roi_info=None, roi_radius=None, roi_center=None, lst=None,
                gradient_mode=None, memsave=False, vmemavail=None, store_prev_skymodel_file=None):
        """One snapshot; same arguments as interferometry.py:5874-5878."""
        nbl, nchan = self.baselines.shape[0], self.channels.size
        if gradient_mode is not None:                                                  # :6306-6311
            if not isinstance(gradient_mode, str):
                raise TypeError("Input gradient_mode must be a string")
            if gradient_mode.lower() not in ["baseline", "skypos", "frequency"]:
                raise ValueError("Invalid value specified in input gradient_mode")
            if gradient_mode.lower() != "baseline":
                raise NotImplementedError("only gradient_mode='baseline' is computed (as in the reference, :6312)")
            if self.gradient_mode is None:
                self.gradient_mode = gradient_mode
        bandpass = NP.asarray(bandpass)
        if bandpass.ndim == 1:                                                         # :5993-5996
            if bandpass.size != nchan:
                raise ValueError("Specified bandpass incompatible with the number of frequency channels")
            bp_t = engine._f64(bandpass, self.device)
        elif bandpass.ndim in (2, 3):                                                  # :6002-6022
            if bandpass.shape[1] != nchan:
                raise ValueError("Specified bandpass incompatible with the number of frequency channels")
            if bandpass.shape[0] != nbl:
                raise ValueError("Specified bandpass incompatible with the number of interferometers")
            if bandpass.ndim == 3:
                if bandpass.shape[2] != 1:
                    raise ValueError("Bandpass can have only one layer for this instance of accumulation.")
                bandpass = bandpass[:, :, 0]
            bp_t = self._compact(bandpass)
        else:
            raise ValueError("Specified bandpass has incompatible dimensions")

        if not isinstance(Tsysinfo, dict):                                             # :6026-6039
            raise TypeError("Input Tsysinfo must be a dictionary")
        Tsys = None
        if Tsysinfo.get("Tnet", None) is not None:
            Tsys = Tsysinfo["Tnet"]
        if Tsys is None:
            try:
                Tsys = Tsysinfo["Trx"] + Tsysinfo["Tant"]["T0"] * (self.channels / Tsysinfo["Tant"]["f0"]) ** Tsysinfo["Tant"]["spindex"]
            except KeyError:
                raise KeyError("One or more keys not found in input Tsysinfo")
            Tsys = NP.asarray(Tsys, dtype=NP.float64).reshape(1, -1)
        if bpcorrect is not None:                                                      # :6042-6053
            if not isinstance(bpcorrect, NP.ndarray):
                raise TypeError("Input specifying bandpass correction must be a numpy array")
            if bpcorrect.size == nchan:
                bpcorrect = bpcorrect.reshape(1, -1)
            elif bpcorrect.size == nbl:
                bpcorrect = bpcorrect.reshape(-1, 1)
            elif bpcorrect.size == nbl * nchan:
                bpcorrect = bpcorrect.reshape(-1, nchan)
            else:
                raise ValueError("Input bpcorrect has dimensions incompatible with the number of baselines and frequencies")
            Tsys = NP.asarray(Tsys, dtype=NP.float64) * bpcorrect
        if isinstance(Tsys, (int, float)):                                             # :6055-6063
            if Tsys < 0.0:
                raise ValueError("Tsys found to be negative.")
            Tsys_t = engine._f64(NP.full(nchan, float(Tsys)), self.device)
        elif isinstance(Tsys, (list, tuple, NP.ndarray)):                              # :6064-6084
            Tsys = NP.asarray(Tsys, dtype=NP.float64)
            if NP.any(Tsys < 0.0):
                raise ValueError("Tsys should be non-negative.")
            if Tsys.size == nbl:                                                       # the reference tests nbl first (:6068)
                Tsys_t = self._compact(NP.repeat(Tsys.reshape(-1, 1), nchan, axis=1))
            elif Tsys.size == nchan:
                Tsys_t = engine._f64(Tsys.ravel(), self.device)
            elif Tsys.size == nbl * nchan:
                Tsys_t = self._compact(Tsys.reshape(-1, nchan))
            else:
                raise ValueError("Specified Tsys has incompatible dimensions with the number of baselines and/or number of frequency channels.")
        else:
            raise TypeError("Tsys should be a scalar, list, tuple, or numpy array")

        if hasattr(timeobj, "sidereal_time"):                                          # :6113
            lst = timeobj.sidereal_time("apparent").deg
        if lst is not None:
            lst = float(NP.asarray(lst).ravel()[0])
        pointing_center = NP.asarray(pointing_center, dtype=NP.float64).reshape(1, -1)
        pc = pointing_center[0]

        # pointing centre -> Alt-Az -> direction cosines (:6155-6164)
        if self.pointing_coords == "hadec":
            pc_altaz = GEOM.hadec2altaz(pc, self.latitude, units="degrees")[0]
        elif self.pointing_coords == "radec":
            if lst is None:
                raise ValueError("LST must be provided. Sky coordinates are in Alt-Az format while pointing center is in RA-Dec format.")
            pc_altaz = GEOM.hadec2altaz(NP.asarray([lst - pc[0], pc[1]]), self.latitude, units="degrees")[0]
        else:
            pc_altaz = pc
        pc_dircos = GEOM.altaz2dircos(pc_altaz, "degrees")[0]

        if self._d_bl is None:                                                         # :6151-6153
            bl_local = self.baselines
            if self.baseline_coords == "equatorial":
                bl_local = GEOM.xyz2enu(self.baselines, self.latitude, "degrees")
            self._d_bl = engine._f64(bl_local, self.device)

        for attr in ("location",):
            if not hasattr(skymodel, attr):
                raise TypeError("skymodel should be an instance of class SkyModel.")    # :6171
        sky = self._sky_to_device(skymodel)

        # sky positions in the frame the cull kernel understands (:6174-6180)
        if self.skycoords == "radec":
            if lst is None:
                raise ValueError("LST must be provided to observe a sky model in RA-Dec coordinates.")
            skypos = torch.stack((lst - sky["location"][:, 0], sky["location"][:, 1]), dim=1).contiguous()
            coords = "hadec"
        else:
            skypos, coords = sky["location"], self.skycoords

        pbeam = None
        if roi_info is not None:                                                       # :6189-6202
            if ("ind" not in roi_info) or ("pbeam" not in roi_info):
                raise KeyError('Both "ind" and "pbeam" keys must be present in dictionary roi_info')
        if (roi_info is not None) and (roi_info["ind"] is not None) and (roi_info["pbeam"] is not None):
            m2 = NP.asarray(roi_info["ind"]).ravel()
            if m2.size > 0:
                try:
                    pb_host = NP.asarray(roi_info["pbeam"], dtype=NP.float64).reshape(-1, nchan)
                except ValueError:
                    raise ValueError('Number of columns of primary beam in key "pbeam" of dictionary roi_info must be equal to number of frequency channels.')
                if m2.size != pb_host.shape[0]:
                    raise ValueError("Values in keys ind and pbeam in must carry same number of elements.")
                pbeam = engine._f64(pb_host, self.device)
                sel = torch.as_tensor(m2.astype(NP.int64), device=self._dev_str())
                dircos, _ = engine.sky_cull(skypos.index_select(0, sel).contiguous(), coords, latitude_deg=self.latitude,
                                            roi_radius_deg=180.0, device=self.device)
                index = sel.to(torch.int32)
            else:
                dircos, index = skypos[:0], torch.empty(0, dtype=torch.int32, device=self._dev_str())
        else:
            if roi_radius is None:                                                     # :6204-6216
                roi_radius = 90.0
            if roi_center is None:
                roi_center = "zenith"
            elif roi_center not in ("zenith", "pointing_center"):
                raise ValueError('Center of region of interest, roi_center, must be set to "zenith" or "pointing_center".')
            center = pc_dircos if roi_center == "pointing_center" else None
            dircos, index = engine.sky_cull(skypos, coords, latitude_deg=self.latitude, roi_radius_deg=roi_radius,
                                            roi_center_dircos=center, device=self.device)
        nsrc = int(index.shape[0])

        if nsrc > 0:
            if pbeam is not None:
                beam = engine.make_beam_desc(element=engine._lib.BEAM_TABLE)
            elif isinstance(pb_info, dict) and pb_info.get("external_beam", None) is not None:
                # gridded HEALPix beam gathered on the device (the run_prisim external-beam step, :1897-1908)
                hb = pb_info["external_beam"]
                if hb.nchan != nchan:
                    raise ValueError("external beam was prepared for a different number of channels")
                pbeam, logmax = hb.table(dircos, nsrc)
                beam = engine.make_beam_desc(element=engine._lib.BEAM_LOGTABLE, d_logmax=logmax)
            else:                                                                      # :6251-6252
                beam = PB.beam_desc_from_telescope(self.telescope, pointing_info=pb_info, pointing_center=pc_altaz,
                                                   skyunits="altaz", device=self.device)
            fwhm = None
            if "fwhm" in sky:                                                          # :6258-6267
                fwhm = sky["fwhm"].index_select(0, index.to(torch.int64)).contiguous()
            skyvis, grad = self._phase_sum(dircos, index, nsrc, sky["spec"], beam, pbeam, pc_dircos, fwhm,
                                           gradient=gradient_mode is not None)
            self.obs_catalog_indices = self.obs_catalog_indices + [index.cpu().numpy().astype(NP.int64)]   # :6377
        else:                                                                          # :6378-6382
            warnings.warn("No sources found in the catalog within matching radius. Simply populating the observed visibilities and/or gradients with noise.")
            out, self.next_skyvis_out = self.next_skyvis_out, None
            skyvis = out.zero_() if out is not None else torch.zeros((nbl, nchan), dtype=torch.complex128, device=self._dev_str())
            grad = torch.zeros((3, nbl, nchan), dtype=torch.complex128, device=self._dev_str()) if gradient_mode is not None else None

        # bookkeeping (:6103-6108, :6384-6399)
        if not self.timestamp:
            self.pointing_center = pointing_center.copy()
            self.phase_center = pointing_center.copy()
        else:
            self.pointing_center = NP.vstack((self.pointing_center, pointing_center))
            self.phase_center = NP.vstack((self.phase_center, pointing_center))
        if not skyvis.is_contiguous():                   # strided rows of another rank's buffer (interleaved shard): keep a dense local copy
            skyvis = skyvis.contiguous()
        self._skyvis.append(skyvis)
        if gradient_mode is not None:                                                  # :6386-6394
            self._gradient.append(grad)
        self._bp.append(bp_t)
        self._Tsys.append(Tsys_t)
        self.Tsysinfo += [Tsysinfo]
        self.timestamp = self.timestamp + [timeobj.jd if hasattr(timeobj, "jd") else timeobj]
        self.t_acc = self.t_acc + [t_acc]
        self.t_obs += t_acc
        self.n_acc += 1
        self.lst = self.lst + [lst]

    def _phase_sum(self, dircos, index, nsrc, spec, beam, pbeam, pc_dircos, fwhm, gradient=False):
        """Amplitude table + phase sum with precision control (see `precision` in __init__).  Returns (V, G): G is the
        [3, nbl, nchan] gradient of V w.r.t. the baseline vector (three more phase sums over the amplitude table scaled
        by one direction cosine each, interferometry.py:6338/:6343) or None.  Unlike the reference -- whose gradient
        branch reads direction cosines that only exist when the sky model has src_shape (:6263) -- it is also computed
        for point-source skies.  Each gradient component goes through the same precision control as V."""
        nbl, nchan = self.baselines.shape[0], self.channels.size
        # brightest sources first (the sum does not depend on the order; obs_catalog_indices keeps the catalogue order):
        # the fp32 kernel moves their partial sums to fp64 after every tile (engine.brightness_order)
        nbright = 0
        if self.sort_by_brightness and nsrc > 64:
            perm, nbright = engine.brightness_order(dircos, index, nsrc, spec, beam, self.channels, pbeam=pbeam, device=self.device)
            dircos = dircos.index_select(0, perm).contiguous()
            index = index.index_select(0, perm).contiguous()
            fwhm = None if fwhm is None else fwhm.index_select(0, perm).contiguous()
            pbeam = None if pbeam is None else pbeam.index_select(0, perm).contiguous()
        kw = dict(pbeam=pbeam, device=self.device)
        uniform = engine.channels_uniform(self.channels)      # the library's own criterion (pb200_channels_uniform)
        if self.precision == "fp64" and not uniform:
            raise ValueError("precision='fp64' needs uniformly spaced channels (the fp64 kernel is a channel recurrence); "
                             "this channel grid takes the direct fp32 kernel -- use precision='fp32' or 'auto'")

        def run64(bl, amp64=None, out=None):
            if amp64 is None:
                amp64 = engine.amp_table(dircos, index, nsrc, spec, beam, self.channels, dtype=torch.float64, **kw)
            return engine.skyvis(dircos, amp64, nsrc, bl, pc_dircos, self.channels, src_fwhm_deg=fwhm, method="fp64",
                                 device=self.device, out=out)

        def scaled(amp_any, i):
            return engine.amp_scale(amp_any, nsrc, nchan, dircos, i, device=self.device)

        out, self.next_skyvis_out = self.next_skyvis_out, None
        if out is not None and (tuple(out.shape) != (nbl, nchan) or out.dtype != torch.complex128 or (nchan > 1 and out.stride(1) != 1)):
            raise ValueError("next_skyvis_out must be a [nbl, nchan] complex128 CUDA tensor with unit channel stride")
        if uniform and (self.precision == "fp64" or (self.precision == "auto" and self._fp64_sticky)):
            self.precision_report.append({"fp64_baselines": nbl, "nbl": nbl, "audited": 0, "audit_max_err": 0.0})
            amp64 = engine.amp_table(dircos, index, nsrc, spec, beam, self.channels, dtype=torch.float64, **kw)
            grad = torch.stack([run64(self._d_bl, scaled(amp64, i)) for i in range(3)]) if gradient else None
            return run64(self._d_bl, amp64, out=out), grad
        amp = engine.amp_table(dircos, index, nsrc, spec, beam, self.channels, **kw)
        amp64_cache = {}

        def amp64_table():
            if "t" not in amp64_cache:
                amp64_cache["t"] = engine.amp_table(dircos, index, nsrc, spec, beam, self.channels, dtype=torch.float64, **kw)
            return amp64_cache["t"]

        def certified(amp32, amp64_fn, report, out=None):
            """fp32 phase sum of one amplitude table + the 'auto' cancellation test and fp64 audit."""
            skyvis = engine.skyvis(dircos, amp32, nsrc, self._d_bl, pc_dircos, self.channels, src_fwhm_deg=fwhm, device=self.device,
                                   method=self.skyvis_method, out=out, nsrc_bright=nbright)
            if not (self.precision == "auto" and uniform):
                return skyvis
            # (1) cancellation test.  The fp32 kernel's absolute error on incoherent (point-source) skies is
            # ~1-3.5e-6 of the incoherent norm sqrt(mean_f sum_s a^2) (measured), the tolerance 1e-5 of each
            # baseline's rms: baselines whose spectrum cancels below cancel_ratio x that norm go to fp64.
            a2 = torch.linalg.vector_norm(amp32, dtype=torch.float64) / (nchan ** 0.5)       # one pass, no fp64 copy of the table
            rms_b = torch.sqrt(skyvis.real.square().mean(dim=1) + skyvis.imag.square().mean(dim=1))
            low = rms_b < self.cancel_ratio * a2
            flagged = torch.nonzero(low).flatten()
            nflag = int(flagged.numel())
            # (2) sampled fp64 audit.  On coherent skies (smooth diffuse emission: same-sign amplitudes, slowly
            # varying phases) fp32 partial sums are much larger than the incoherent norm and so is the error;
            # a few un-flagged baselines (shortest, longest, random) are recomputed in fp64 and compared.
            # Flagged and audited baselines share ONE fp64 launch (both are far fewer than a wave of CTAs).
            keep = torch.nonzero(~low).flatten()
            sel = keep[:0]
            if keep.numel() > 0 and self.audit_baselines > 0:
                gen = torch.Generator(device="cpu").manual_seed(20261017 + getattr(self, "_drained", 0) + len(self._skyvis))
                pick = torch.randperm(int(keep.numel()), generator=gen)[: max(self.audit_baselines - 2, 0)]
                sel = torch.unique(torch.cat((keep[[0, -1]], keep[pick.to(keep.device)])))
            audit_err, audited = 0.0, int(sel.numel())
            both = torch.cat((flagged, sel))
            if both.numel() > 0:
                amp64 = amp64_fn()
                ref = run64(self._d_bl.index_select(0, both).contiguous(), amp64)
                if audited:
                    ref_a, got = ref[nflag:], skyvis.index_select(0, sel)
                    rms_ref = torch.sqrt(ref_a.real.square().mean(dim=1) + ref_a.imag.square().mean(dim=1)).clamp_min(1e-300)
                    audit_err = float(((got - ref_a).abs().amax(dim=1) / rms_ref).max().item())
                skyvis.index_copy_(0, both, ref)
                if audit_err > self.audit_tolerance:          # fp32 is not good enough on this sky: everything in fp64
                    skyvis.index_copy_(0, keep, run64(self._d_bl.index_select(0, keep).contiguous(), amp64))
                    nflag = nbl
            if report:
                self.precision_report.append({"fp64_baselines": nflag, "nbl": nbl, "audited": audited, "audit_max_err": audit_err})
                self._fp64_sticky = nflag > 0.5 * nbl
            return skyvis

        skyvis = certified(amp, amp64_table, True, out=out)
        grad = None
        if gradient:      # every component is certified like V itself (the l and m weights change sign over the sky: they cancel more)
            grad = torch.stack([certified(scaled(amp, i), lambda i=i: scaled(amp64_table(), i), False) for i in range(3)])
        return skyvis, grad

    # ------------------------------------------------------------------ observing_run
    def observing_run(self, pointing_init, skymodel, t_acc, duration, channels, bpass, Tsys, lst_init, roi_radius=None,
                      roi_center=None, mode="track", pointing_coords=None, freq_scale=None, brightness_units=None,
                      verbose=True, memsave=False, jd_init=2451545.0):
        """Same call as interferometry.py:6414-6417.  The reference's loop (:6641-6647) passes a
        string timestamp and an ndarray Tsys to ``observe`` and no longer runs (Appendix C #1); this
        implements the documented behaviour: an LST ladder (:6607) and a tracking or drifting
        pointing centre (:6609-6633), one ``observe`` per accumulation."""
        if verbose:
            print("Preparing an observing run...")
        if not isinstance(t_acc, (int, float)) or t_acc <= 0.0:
            raise ValueError("Accumulation interval must be a positive scalar")
        if not isinstance(duration, (int, float)) or duration <= 0.0:
            raise ValueError("Observing duration must be a positive scalar")
        if duration < t_acc:
            duration = t_acc
        n_acc = int(duration / t_acc)
        nbl, nchan = self.baselines.shape[0], self.channels.size
        bpass = NP.asarray(bpass, dtype=NP.float64)
        if bpass.size == nchan:                                                        # :6560-6571
            bpass = bpass.reshape(1, nchan, 1)
        elif bpass.size == nbl * nchan:
            bpass = bpass.reshape(nbl, nchan, 1)
        elif bpass.size == nbl * nchan * n_acc:
            bpass = bpass.reshape(nbl, nchan, n_acc)
        else:
            raise ValueError("Dimensions of bpass incompatible with the number of frequency channels, baselines and number of accumulations.")
        if not isinstance(Tsys, (int, float, list, tuple, NP.ndarray)):                # :6573-6576
            raise TypeError("Tsys must be a scalar, list, tuple or numpy array")
        Tsys = NP.asarray(Tsys, dtype=NP.float64).reshape(-1)
        if Tsys.size == 1:                                                             # :6578-6599
            Tsys = Tsys[0] + NP.zeros((1, nchan, 1))
        elif Tsys.size == nchan:
            Tsys = Tsys.reshape(1, nchan, 1)
        elif Tsys.size == nbl:
            Tsys = NP.repeat(Tsys.reshape(nbl, 1, 1), nchan, axis=1)
        elif Tsys.size == nbl * nchan:
            Tsys = Tsys.reshape(nbl, nchan, 1)
        elif Tsys.size == nbl * nchan * n_acc:
            Tsys = Tsys.reshape(nbl, nchan, n_acc)
        else:
            raise ValueError("Dimensions of Tsys incompatible with the number of frequency channels, baselines and number of accumulations.")
        if not isinstance(lst_init, (int, float)):
            raise TypeError("Starting LST should be a scalar")
        lst = (lst_init + (t_acc / 3.6e3) * NP.arange(n_acc)) * 15.0                   # :6607, degrees
        pointing_init = NP.asarray(pointing_init, dtype=NP.float64).ravel()
        lst0_deg = lst_init * 15.0
        if mode == "track":                                                            # :6611-6621
            if pointing_coords == "hadec":
                pointing = NP.asarray([lst0_deg - pointing_init[0], pointing_init[1]])
            elif (pointing_coords == "radec") or (pointing_coords is None):
                pointing = pointing_init
            elif pointing_coords == "altaz":
                hadec = GEOM.altaz2hadec(pointing_init, self.latitude, units="degrees")[0]
                pointing = NP.asarray([lst0_deg - hadec[0], hadec[1]])
            else:
                raise ValueError('pointing_coords can only be set to "hadec", "radec" or "altaz".')
            self.pointing_coords = "radec"
            self.phase_center_coords = "radec"
        elif mode == "drift":                                                          # :6622-6633
            if pointing_coords == "radec":
                pointing = NP.asarray([lst0_deg - pointing_init[0], pointing_init[1]])
            elif (pointing_coords == "hadec") or (pointing_coords is None):
                pointing = pointing_init
            elif pointing_coords == "altaz":
                pointing = GEOM.altaz2hadec(pointing_init, self.latitude, units="degrees")[0]
            else:
                raise ValueError('pointing_coords can only be set to "hadec", "radec" or "altaz".')
            self.pointing_coords = "hadec"
            self.phase_center_coords = "hadec"
        else:
            raise ValueError('mode must be "track" or "drift"')
        for i in range(n_acc):                                                         # :6641-6647
            timeobj = SimpleTime(jd_init + i * t_acc / 86400.0, lst[i] % 360.0)
            bp_i = bpass[:, :, i % bpass.shape[2]]
            Ts_i = Tsys[:, :, i % Tsys.shape[2]]
            self.observe(timeobj, {"Tnet": Ts_i[0] if Ts_i.shape[0] == 1 else Ts_i},
                         bp_i[0] if bp_i.shape[0] == 1 else bp_i, pointing, skymodel, t_acc,
                         brightness_units=brightness_units, roi_radius=roi_radius, roi_center=roi_center,
                         lst=lst[i] % 360.0, memsave=memsave)
        self.t_obs = duration                                                          # :6654-6655
        self.n_acc = n_acc
        if verbose:
            print("Observing run completed successfully.")

    # ------------------------------------------------------------------ noise
    def _aeff_effq(self):
        if self._d_aeff is None:
            self._d_aeff = self._compact(self.A_eff)
            self._d_effq = self._compact(self.eff_Q)
        return self._d_aeff, self._d_effq

    def generate_noise(self):
        """interferometry.py:6661-6693: thermal rms and a noise realisation for every snapshot."""
        if self.flux_unit.upper() not in ("JY", "K"):
            raise ValueError("Flux density units can only be in Jy or K.")
        aeff, effq = self._aeff_effq()
        nbl, nchan = self.baselines.shape[0], self.channels.size
        self._rms, self._noise = [], []
        base = getattr(self, "_drained", 0)
        seed = _realisation_seed(self.noise_seed, self._noise_realisation)        # a new realisation on every call (:6693)
        self._noise_realisation += 1
        for t in range(len(self._skyvis)):
            rms, nz, _ = engine.noise(None, self._Tsys[t], aeff, effq, self.freq_resolution, self.t_acc[base + t],
                                      seed, nbl, nchan, snapshot=base + t, bl_offset=self.bl_offset, bl_step=self.bl_step,
                                      nbl_total=self.nbl_total, flux_unit_k=(self.flux_unit.upper() == "K"),
                                      want=("rms", "noise"), device=self.device)
            self._rms.append(rms)
            self._noise.append(nz)

    def add_noise(self):
        """interferometry.py:6697-6722 with unity gains (no gain table on the hot path)."""
        if not self._noise:
            raise TypeError("vis_noise_freq has not been generated; call generate_noise() first")
        warnings.warn("Gain table absent. Proceeding with default unity gains")
        self._vis = [engine.add_noise(self._skyvis[t], self._noise[t]) for t in range(len(self._skyvis))]

    # ------------------------------------------------------------------ delay transform
    def _freq_wts_getter(self, freq_wts):
        """Broadcast rules of interferometry.py:8096-8106, kept compact on the device."""
        nbl, nchan, n_acc = self.baselines.shape[0], self.channels.size, self.n_acc
        freq_wts = NP.asarray(freq_wts, dtype=NP.float64)
        if freq_wts.size == nchan:
            w = engine._f64(freq_wts.ravel(), self.device)
            return lambda t: w
        if freq_wts.size == nchan * n_acc:
            w = engine._f64(freq_wts.reshape(nchan, -1).T, self.device)      # [n_acc, nchan]
            return lambda t: w[t]
        if freq_wts.size == nchan * nbl:
            w = engine._f64(freq_wts.reshape(-1, nchan), self.device)
            return lambda t: w
        if freq_wts.size == nchan * nbl * n_acc:
            w = engine._f64(NP.ascontiguousarray(NP.moveaxis(freq_wts.reshape(nbl, nchan, n_acc), 2, 0)), self.device)
            return lambda t: w[t]
        raise ValueError("window shape dimensions incompatible with number of channels and/or number of tiemstamps.")

    def delay_transform(self, pad=1.0, freq_wts=None, verbose=True):
        """interferometry.py:8052-8137.  Transforms whichever of skyvis / vis / noise exist
        (the reference raises TypeError when noise is missing, Appendix C #6)."""
        if verbose:
            print("Preparing to compute delay transform...\n\tChecking input parameters for compatibility...")
        if not isinstance(pad, (int, float)):
            raise TypeError("pad fraction must be a scalar value.")
        if pad < 0.0:
            pad = 0.0
            if verbose:
                warnings.warn("\tPad fraction found to be negative. Resetting to 0.0 (no padding will be applied).")
        if freq_wts is not None:
            self._bp_wts = self._freq_wts_getter(freq_wts)
        nbl, nchan = self.baselines.shape[0], self.channels.size
        self.lags = NP.fft.fftshift(NP.fft.fftfreq(nchan, d=self.freq_resolution))    # :8114
        products = {"skyvis": self._skyvis, "vis": self._vis, "noise": self._noise}
        self._lag = {"skyvis": [], "vis": [], "noise": [], "kernel": []}
        base = getattr(self, "_drained", 0)
        for t in range(len(self._skyvis)):
            bp = self._bp[t]
            wts = None if self._bp_wts is None else self._bp_wts(base + t)
            for name, lst in products.items():
                if lst:
                    self._lag[name].append(engine.delay_transform(lst[t], bp, wts, self.freq_resolution, pad=pad,
                                                                  downsample=True))
            krows = nbl if (bp.ndim == 2 or (wts is not None and wts.ndim == 2)) else 1
            kern = engine.delay_transform(None, bp, wts, self.freq_resolution, pad=pad, downsample=True, nrows=krows,
                                          nchan=nchan, device=self.device)
            self._lag["kernel"].append(kern if krows == nbl else kern[0])
        if verbose:
            print("delay_transform() completed successfully.")

    # ------------------------------------------------------------------ sub-band delay transforms
    def subband_windows(self, bw_eff, freq_center=None, shape=None):
        """The [nwin, nchan] sub-band weights multi_window_delay_transform builds (interferometry.py:8199-8265): for
        every (effective bandwidth, centre frequency) a window of round(bw_eff / (w_frac df)) samples of `shape`
        ('rect' | 'bhw' | 'bnw'; w_frac = its equivalent-rectangle width fraction) centred on the nearest channel,
        clipped to the band, ordered by centre channel."""
        if not isinstance(bw_eff, (int, float, list, NP.ndarray)):
            raise TypeError("Effective bandwidth must be a scalar, list or numpy array")
        bw = NP.asarray(bw_eff, dtype=NP.float64).reshape(-1)
        if NP.any(bw <= 0.0):
            raise ValueError("All values in effective bandwidth must be strictly positive")
        f, df = self.channels, self.freq_resolution
        if freq_center is None:
            fc = NP.asarray(f[int(0.5 * f.size)]).reshape(-1)
        elif isinstance(freq_center, (int, float, list, NP.ndarray)):
            fc = NP.asarray(freq_center, dtype=NP.float64).reshape(-1)
            if NP.any((fc <= f.min()) | (fc >= f.max())):
                raise ValueError("Frequency centers must lie strictly inside the observing band")
        else:
            raise TypeError("Frequency center(s) must be scalar, list or numpy array")
        if bw.size == 1 and fc.size > 1:
            bw = NP.repeat(bw, fc.size)
        elif bw.size > 1 and fc.size == 1:
            fc = NP.repeat(fc, bw.size)
        elif bw.size != fc.size:
            raise ValueError("Effective bandwidth(s) and frequency center(s) must have same number of elements")
        if shape is not None:
            if not isinstance(shape, str):
                raise TypeError("Window shape must be a string")
            if shape not in ["rect", "bhw", "bnw", "RECT", "BHW", "BNW"]:
                raise ValueError("Invalid value for window shape specified.")
        else:
            shape = "rect"
        from .delay_spectrum import windowing, window_N2width
        nwin = NP.round(bw / window_N2width(shape=shape) / df).astype(int)              # :8236-8238
        centre = NP.rint((fc - f[0]) / df).astype(int)                                  # nearest channel, :8246
        order = NP.argsort(centre, kind="stable")
        wts = NP.zeros((fc.size, f.size))
        for row, (c, n) in enumerate(zip(centre[order], nwin[order])):
            w = windowing(int(n), shape=shape.lower(), centering=True)
            k = c + NP.arange(int(n)) - int(n / 2)                                      # :8255
            inside = (k >= 0) & (k < f.size)
            wts[row, k[inside]] = w[inside]
        return wts

    def multi_window_delay_transform(self, bw_eff, freq_center=None, shape=None, pad=1.0, verbose=True):
        """Delay transforms over several sub-bands, same call and products as interferometry.py:8141-8287: returns
        {'skyvis_lag', 'vis_noise_lag', 'lag_kernel'} of shape [nbl, nwin, nchan, nsnap] and 'lag_corr_length' [nwin].
        Each (window, snapshot) is one launch of the delay-transform kernel with the window as its weights."""
        wts_host = self.subband_windows(bw_eff, freq_center=freq_center, shape=shape)
        if not isinstance(pad, (int, float)):
            raise TypeError("pad fraction must be a scalar value.")
        if pad < 0.0:
            pad = 0.0
            if verbose:
                warnings.warn("\tPad fraction found to be negative. Resetting to 0.0 (no padding will be applied).")
        nbl, nchan, nsnap = self.baselines.shape[0], self.channels.size, len(self._skyvis)
        wts_dev = engine._f64(wts_host, self.device)
        nout = engine.delay_nout(nchan, pad, True)
        out = {}
        for key, lst in (("skyvis_lag", self._skyvis), ("vis_noise_lag", self._noise), ("lag_kernel", None)):
            if lst is not None and not lst:
                continue                                                                 # product not generated (yet)
            res = torch.empty((nbl, wts_host.shape[0], nout, nsnap), dtype=torch.complex128, device=self._dev_str())
            for t in range(nsnap):
                for i in range(wts_host.shape[0]):
                    if lst is None:
                        kern = engine.delay_transform(None, self._bp[t], wts_dev[i], self.freq_resolution, pad=pad, downsample=True,
                                                      nrows=nbl if self._bp[t].ndim == 2 else 1, nchan=nchan, device=self.device)
                        res[:, i, :, t] = kern
                    else:
                        res[:, i, :, t] = engine.delay_transform(lst[t], self._bp[t], wts_dev[i], self.freq_resolution, pad=pad,
                                                                 downsample=True)
            out[key] = res.cpu().numpy()
        out["lag_corr_length"] = nchan / NP.sum(wts_host, axis=1)                        # :8287
        if verbose:
            print("multi_window_delay_transform() completed successfully.")
        return out

    # ------------------------------------------------------------------ redundant baselines
    def duplicate_measurements(self, blgroups=None):
        """Expand the simulated (unique) baselines into their redundant sets, same call as interferometry.py:6823-6907:
        every per-baseline product is repeated ``len(group)`` times along the baseline axis (one device row gather per
        snapshot), labels / baselines / lengths / projected baselines / per-baseline A_eff, eff_Q, Tsys, bandpass
        follow, and noise is regenerated for the expanded set (:6905-6906).  ``blgroups``: {label tuple: sequence of
        label tuples redundant with it}; a key missing from its own group is prepended, labels without a group are
        kept once, a label in two groups raises ValueError.  Nothing happens when the groups hold no more baselines
        than are present (:6852-6857)."""
        if blgroups is None:
            blgroups = self.blgroups
        if not isinstance(blgroups, dict):
            raise TypeError("Input blgroups must be a dictionary")
        labels = [tuple(l) for l in (self.labels.tolist() if isinstance(self.labels, NP.ndarray) else self.labels)]
        groups = {tuple(k): [tuple(l) for l in (v.tolist() if isinstance(v, NP.ndarray) else v)] for k, v in blgroups.items()}
        nbl_new = len(self.bl_reversemap) if self.bl_reversemap is not None else sum(len(v) for v in groups.values())
        if len(labels) >= nbl_new:
            return
        for key in list(groups):
            use = key
            if key not in labels:
                if tuple(reversed(key)) not in labels:
                    raise KeyError("Input label {0} not found in attribute labels".format(key))
                use = tuple(reversed(key))
                groups.setdefault(use, groups[key])
            if use not in groups[use]:
                groups[use] = [use] + groups[use]
        num_list, out = [], []
        for label in labels:
            if label in groups:
                num_list.append(len(groups[label]))
                for lbl in groups[label]:
                    if lbl in out:
                        raise ValueError("Label {0} repeated in more than one baseline group".format(lbl))
                    out.append(lbl)
            else:
                num_list.append(1)
                out.append(label)
        num = NP.asarray(num_list, dtype=NP.int64)
        nbl_old = len(labels)
        idx = torch.repeat_interleave(torch.arange(nbl_old, device=self._dev_str()), torch.as_tensor(num, device=self._dev_str()))
        rep = lambda t: t.index_select(0, idx) if (t.dim() == 2 and t.shape[0] == nbl_old) else t      # [nchan] rows are shared
        self._skyvis = [t.index_select(0, idx) for t in self._skyvis]                    # :6887-6888
        self._gradient = [t.index_select(1, idx) for t in self._gradient]                # :6889-6890
        self._bp = [rep(t) for t in self._bp]
        self._Tsys = [rep(t) for t in self._Tsys]
        if isinstance(self.labels, NP.ndarray) and self.labels.dtype.names:
            self.labels = NP.asarray(out, dtype=self.labels.dtype)                       # :6892
        else:
            self.labels = out
        self.baselines = NP.repeat(self.baselines, num, axis=0)
        if self.projected_baselines is not None:
            self.projected_baselines = NP.repeat(self.projected_baselines, num, axis=0)
        self.baseline_lengths = NP.repeat(self.baseline_lengths, num)
        self.baseline_orientations = NP.repeat(self.baseline_orientations, num)
        self.eff_Q = NP.repeat(self.eff_Q, num, axis=0)
        self.A_eff = NP.repeat(self.A_eff, num, axis=0)
        if self._d_bl is not None:
            self._d_bl = self._d_bl.index_select(0, idx).contiguous()
        self._d_aeff = self._d_effq = None
        self._lag = {}
        self._vis, self._noise, self._rms = [], [], []
        self.nbl_total = self.baselines.shape[0] if self.nbl_total == nbl_old else self.nbl_total
        self.generate_noise()                                                            # :6905-6906
        self.add_noise()

    # ------------------------------------------------------------------ bounded-memory streaming
    def drain(self, sink, noise=True, delay_transform=None, ring=0):
        """Stream the resident snapshots out and free their device memory (a 1000-snapshot HERA-350 run is
        ~7.5 GB per snapshot with all products and cannot stay resident; the reference keeps growing numpy
        arrays, interferometry.py:6390).  For every resident snapshot: optionally generate noise + add it
        (same Philox stream as generate_noise/add_noise: keyed by the global snapshot index) and delay-transform
        (``delay_transform=dict(pad=..., freq_wts=...)``), then call ``sink(j, products)`` with a dict of
        [nbl, n] CUDA tensors ('skyvis_freq', 'vis_freq', 'vis_noise_freq', 'vis_rms_freq', 'skyvis_lag', ...),
        and drop them.  Bookkeeping (timestamp, lst, t_acc, pointing centres, indices) is kept; snapshot numbering
        continues.

        ``ring=n`` (n >= 1): the noise / delay-transform products are written into n preallocated buffer sets owned by
        this object and used in turn (global snapshot j uses set j % n) instead of being allocated per snapshot: what
        the sink receives stays valid until n more snapshots have been drained -- long enough for an asynchronous
        device->host copy that overlaps the next snapshot -- and a long run makes no allocator calls at all."""
        base = getattr(self, "_drained", 0)
        nres = len(self._skyvis)
        if nres == 0:
            return 0
        if noise:
            aeff, effq = self._aeff_effq()
        nbl, nchan = self.baselines.shape[0], self.channels.size
        getter = None
        if delay_transform is not None and delay_transform.get("freq_wts", None) is not None:
            getter = self._freq_wts_getter(delay_transform["freq_wts"])
        pad = 1.0 if delay_transform is None else delay_transform.get("pad", 1.0)
        nout = engine.delay_nout(nchan, pad, True) if delay_transform is not None else 0

        def pooled(slot, name, shape, dtype):
            pool = self.__dict__.setdefault("_drain_pool", {})
            t = pool.get((slot, name))
            if t is None or tuple(t.shape) != tuple(shape):
                t = pool[(slot, name)] = torch.empty(shape, dtype=dtype, device=self._dev_str())
            return t

        for t in range(nres):
            slot = (base + t) % ring if ring else None
            buf = (lambda name, shape, dtype=torch.complex128: pooled(slot, name, shape, dtype)) if ring else (lambda *a, **k: None)
            prod = {"skyvis_freq": self._skyvis[t]}
            if self._gradient:
                prod["gradient_" + self.gradient_mode] = self._gradient[t]
            if noise:
                outs = {"rms": buf("rms", (nbl, nchan), torch.float64), "noise": buf("noise", (nbl, nchan))} if ring else None
                rms, nz, _ = engine.noise(None, self._Tsys[t], aeff, effq, self.freq_resolution, self.t_acc[base + t],
                                          self.noise_seed, nbl, nchan, snapshot=base + t, bl_offset=self.bl_offset, bl_step=self.bl_step,
                                          nbl_total=self.nbl_total, flux_unit_k=(self.flux_unit.upper() == "K"),
                                          want=("rms", "noise"), device=self.device, out=outs)
                prod.update(vis_rms_freq=rms, vis_noise_freq=nz, vis_freq=engine.add_noise(self._skyvis[t], nz, out=buf("vis", (nbl, nchan))))
            if delay_transform is not None:
                wts = None if getter is None else getter(base + t)          # per-snapshot weights are indexed by the global snapshot
                for key in [k for k in ("skyvis_freq", "vis_freq", "vis_noise_freq") if k in prod]:
                    prod[key.replace("_freq", "_lag")] = engine.delay_transform(prod[key], self._bp[t], wts, self.freq_resolution,
                                                                               pad=pad, downsample=True, out=buf(key + "_lag", (nbl, nout)))
            sink(base + t, prod)
        self._skyvis, self._vis, self._noise, self._rms, self._bp, self._Tsys, self._gradient = [], [], [], [], [], [], []
        self._lag = {}
        self._drained = base + nres
        return nres

    # ------------------------------------------------------------------ phase centring / uvw
    def _centre_to_dircos(self, centre, coords):
        """[n,2|3] phase centres in `coords` -> ENU direction cosines, per snapshot (uses self.lst)."""
        centre = NP.array(centre, dtype=NP.float64)
        lst = NP.asarray(self.lst, dtype=NP.float64)
        if coords == "radec":
            centre = NP.stack((lst - centre[:, 0], centre[:, 1]), axis=1)
            coords = "hadec"
        if coords == "hadec":
            centre = GEOM.hadec2altaz(centre, self.latitude, units="degrees")
            coords = "altaz"
        if coords == "altaz":
            centre = GEOM.altaz2dircos(centre, units="degrees")
        return centre

    def phase_centering(self, ref_point, do_delay_transform=False, verbose=True):
        """interferometry.py:7712-7884: multiply skyvis / vis / noise by exp(-2 pi i f b.(s_old - s_new)/c)
        and update ``phase_center`` / ``phase_center_coords`` (kept in the current coordinate system)."""
        if ref_point is None:
            raise ValueError("Invalid input specified in ref_point")
        if not isinstance(ref_point, dict):
            raise TypeError("Input ref_point must be a dictionary")
        if ("location" not in ref_point) or ("coords" not in ref_point):
            raise KeyError('Both keys "location" and "coords" must be specified in input dictionary ref_point')
        phase_center, coords_new = ref_point["location"], ref_point["coords"]
        nsnap = len(self.lst)
        if phase_center is None:
            raise ValueError("Valid phase center not specified in input ref_point")
        if not isinstance(phase_center, NP.ndarray):
            raise TypeError("Phase center must be a numpy array")
        phase_center = phase_center.reshape(-1, phase_center.shape[-1]).astype(NP.float64)
        if phase_center.shape[0] == 1:
            phase_center = NP.repeat(phase_center, nsnap, axis=0)
        elif phase_center.shape[0] != nsnap:
            raise ValueError("One phase center must be provided for every timestamp.")
        if coords_new not in ("dircos", "altaz", "hadec", "radec"):
            raise ValueError("Invalid phase center coordinate system specified")
        if coords_new == "dircos":
            if (phase_center.shape[1] < 2) or (phase_center.shape[1] > 3):
                raise ValueError("Dimensions incompatible for direction cosine positions")
            if NP.any(NP.sqrt(NP.sum(phase_center ** 2, axis=1)) > 1.0):
                raise ValueError("direction cosines found to be exceeding unit magnitude.")
            if phase_center.shape[1] == 2:
                phase_center = NP.hstack((phase_center, (1.0 - NP.sqrt(NP.sum(phase_center ** 2, axis=1))).reshape(-1, 1)))  # :7787 (sic)
        new_dircos = self._centre_to_dircos(phase_center, coords_new)                       # :7813-7846
        cur_dircos = self._centre_to_dircos(self.phase_center, self.phase_center_coords)    # :7853-7863
        # the stored phase centre stays in the array's current coordinate system (:7778-7846, :7867-7868)
        lst = NP.asarray(self.lst, dtype=NP.float64)
        cur = self.phase_center_coords
        if cur == "altaz":
            stored = GEOM.dircos2altaz(new_dircos, units="degrees")
        else:
            hadec = GEOM.altaz2hadec(GEOM.dircos2altaz(new_dircos, units="degrees"), self.latitude, units="degrees")
            if coords_new in ("hadec", "radec") and cur in ("hadec", "radec"):                # avoid a round trip through alt-az
                hadec = phase_center.copy() if coords_new == "hadec" else NP.stack((lst - phase_center[:, 0], phase_center[:, 1]), axis=1)
            stored = hadec if cur == "hadec" else NP.stack((lst - hadec[:, 0], hadec[:, 1]), axis=1)
        pos_diff = cur_dircos - new_dircos                                                  # :7866
        base = getattr(self, "_drained", 0)                 # snapshots already streamed out by drain() are not resident
        for t in range(len(self._skyvis)):
            for lst_t in (self._skyvis, self._vis, self._noise):                            # :7871-7881
                if lst_t:
                    engine.phase_rotate(lst_t[t], self._d_bl, pos_diff[base + t], self.channels)
        self.phase_center = stored
        if do_delay_transform:
            self.delay_transform(verbose=verbose)

    def project_baselines(self, ref_point):
        """interferometry.py:7888-7995: baselines projected on the (u,v,w) frame of a reference direction,
        [nbl, 3, n_acc].  O(nbl x n_acc) host arithmetic."""
        if ref_point is None:
            raise ValueError("Invalid input specified in ref_point")
        if not isinstance(ref_point, dict):
            raise TypeError("Input ref_point must be a dictionary")
        if ("location" not in ref_point) or ("coords" not in ref_point):
            raise KeyError('Both keys "location" and "coords" must be specified in input dictionary ref_point')
        phase_center, coords = ref_point["location"], ref_point["coords"]
        if not isinstance(phase_center, NP.ndarray):
            raise TypeError("The specified reference point must be a numpy array")
        if not isinstance(coords, str):
            raise TypeError("The specified coordinates of the reference point must be a string")
        if coords not in ["radec", "hadec", "altaz", "dircos"]:
            raise ValueError("Specified coordinates of reference point invalid")
        if phase_center.ndim == 1:
            phase_center = phase_center.reshape(1, -1)
        if phase_center.ndim > 2:
            raise ValueError("Reference point has invalid dimensions")
        if (phase_center.shape[0] != self.n_acc) and (phase_center.shape[0] != 1):
            raise ValueError("Reference point has dimensions incompatible with the number of timestamps")
        if phase_center.shape[0] == 1:
            phase_center = phase_center + NP.zeros(self.n_acc).reshape(-1, 1)
        if coords in ("radec", "hadec", "altaz") and phase_center.shape[1] != 2:
            raise ValueError("Reference point has invalid dimensions")
        if coords == "radec":
            ha, dec = NP.asarray(self.lst) - phase_center[:, 0], phase_center[:, 1]
        elif coords == "hadec":
            ha, dec = phase_center[:, 0], phase_center[:, 1]
        else:
            if coords == "dircos":
                if (phase_center.shape[1] < 2) or (phase_center.shape[1] > 3):
                    raise ValueError("Reference point has invalid dimensions")
                if NP.any(NP.sqrt(NP.sum(phase_center ** 2, axis=1)) > 1.0):
                    raise ValueError("direction cosines found to be exceeding unit magnitude.")
                if phase_center.shape[1] == 2:
                    phase_center = NP.hstack((phase_center, (1.0 - NP.sqrt(NP.sum(phase_center ** 2, axis=1))).reshape(-1, 1)))
                phase_center = GEOM.dircos2altaz(phase_center, units="degrees")
            hadec = GEOM.altaz2hadec(phase_center, self.latitude, units="degrees")
            ha, dec = hadec[:, 0], hadec[:, 1]
        ha, dec = NP.radians(ha).ravel(), NP.radians(dec).ravel()
        eq_baselines = GEOM.enu2xyz(self.baselines, self.latitude, units="degrees")          # :7976
        rot_matrix = NP.asarray([[NP.sin(ha), NP.cos(ha), NP.zeros(ha.size)],
                                 [-NP.sin(dec) * NP.cos(ha), NP.sin(dec) * NP.sin(ha), NP.cos(dec)],
                                 [NP.cos(dec) * NP.cos(ha), -NP.cos(dec) * NP.sin(ha), NP.sin(dec)]])   # :7977-7979
        self.projected_baselines = NP.dot(eq_baselines, rot_matrix)                          # :7985

    def rotate_visibilities(self, ref_point, do_delay_transform=False, verbose=True):
        """interferometry.py:7655-7708: phase_centering + project_baselines."""
        if ref_point is None:
            raise ValueError("Invalid input specified in ref_point")
        if not isinstance(ref_point, dict):
            raise TypeError("Input ref_point must be a dictionary")
        if ("location" not in ref_point) or ("coords" not in ref_point):
            raise KeyError('Both keys "location" and "coords" must be specified in input dictionary ref_point')
        self.phase_centering(ref_point, do_delay_transform=do_delay_transform, verbose=verbose)
        self.project_baselines(ref_point)

    # ------------------------------------------------------------------ on-disk (NPZ only)
    def save(self, outfile, fmt="NPZ", tabtype="BinTableHDU", npz=True, overwrite=False, uvfits_parms=None, verbose=True):
        """The NPZ product of interferometry.py:8859-8863 (same keys).  The HDF5/FITS layouts
        (:8722-8854) need h5py/astropy, which this image does not have: fmt must be 'NPZ'."""
        if not isinstance(outfile, str):
            raise TypeError("Output filename must be a string")
        if fmt.upper() != "NPZ":
            raise NotImplementedError("only the NPZ product is written here (h5py/astropy are unavailable); pass fmt='NPZ'")
        if uvfits_parms is not None:
            raise NotImplementedError("UVFITS export is outside the hot-path scope")
        if npz:
            if (self._vis) and (self._noise):
                NP.savez_compressed(outfile + ".npz", skyvis_freq=self.skyvis_freq, vis_freq=self.vis_freq,
                                    vis_noise_freq=self.vis_noise_freq, lst=self.lst, freq=self.channels, timestamp=self.timestamp,
                                    bl=self.baselines, bl_length=self.baseline_lengths)
            else:
                NP.savez_compressed(outfile + ".npz", skyvis_freq=self.skyvis_freq, lst=self.lst, freq=self.channels,
                                    timestamp=self.timestamp, bl=self.baselines, bl_length=self.baseline_lengths)
            if verbose:
                print("\tInterferometer array information written successfully to NPZ file on disk:\n\t\t{0}\n".format(outfile + ".npz"))
