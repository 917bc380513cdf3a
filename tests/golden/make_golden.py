#!/usr/bin/env python
"""Generate golden vectors by executing the REFERENCE'S OWN source files under Python 3.

Run in the build container only (`python tests/golden/make_golden.py`); needs /root/reference.
Nothing from the reference is copied into the repo: its modules are read from
/root/reference/prisim/*.py at run time, three Python-2 idioms are patched in memory (one
``print`` statement, ``.iteritems()``, ``xrange``), and stub modules stand in for the third-party
packages that are not installed here:

  astroutils (un-vendored, unpinned: setup.py:61)  -> geometry / DSP_modules / catalog / constants
      stubs restating the semantics listed in SURVEY.md section 8c  [AU-memory]
  astropy, h5py, progressbar, distutils, healpy     -> inert stand-ins (never on the tested path)

The outputs (tests/golden/*.npz) therefore pin this repo's oracle against the reference's real
beam, delay, phase-sum, noise-rms and delay-transform code; what remains unpinned is the behaviour
of the stubbed astroutils helpers themselves.
"""
import os
import re
import sys
import types

import numpy as NP
import scipy.constants as FCNST
from scipy import interpolate

REF = "/root/reference/prisim"
OUT = os.path.dirname(os.path.abspath(__file__))
# `make_golden.py tag1 tag2 ...` rewrites only those files (every case still runs, so the shared random stream --
# and with it every other file's content -- is unchanged); no arguments = all files
ONLY = set(sys.argv[1:])


# ------------------------------------------------------------------------------------------------
# astroutils stubs [AU-memory]
# ------------------------------------------------------------------------------------------------
def _mod(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)
    sys.modules[name] = m
    return m


def altaz2dircos(altaz, units=None):
    altaz = NP.asarray(altaz, dtype=float).reshape(-1, 2)
    if units == "degrees":
        altaz = NP.radians(altaz)
    return NP.stack((NP.cos(altaz[:, 0]) * NP.sin(altaz[:, 1]), NP.cos(altaz[:, 0]) * NP.cos(altaz[:, 1]), NP.sin(altaz[:, 0])), axis=1)


def dircos2altaz(dircos, units=None):
    dircos = NP.asarray(dircos, dtype=float).reshape(-1, 3)
    out = NP.stack((NP.arcsin(NP.clip(dircos[:, 2], -1, 1)), NP.arctan2(dircos[:, 0], dircos[:, 1]) % (2 * NP.pi)), axis=1)
    return NP.degrees(out) if units == "degrees" else out


def hadec2altaz(hadec, latitude, units=None):
    shape1d = NP.asarray(hadec).ndim == 1
    hadec = NP.asarray(hadec, dtype=float).reshape(-1, 2)
    if units == "degrees":
        ha, dec, lat = NP.radians(hadec[:, 0]), NP.radians(hadec[:, 1]), NP.radians(latitude)
    else:
        ha, dec, lat = hadec[:, 0], hadec[:, 1], latitude
    north = NP.sin(dec) * NP.cos(lat) - NP.cos(dec) * NP.cos(ha) * NP.sin(lat)
    east = -NP.cos(dec) * NP.sin(ha)
    up = NP.sin(dec) * NP.sin(lat) + NP.cos(dec) * NP.cos(ha) * NP.cos(lat)
    out = NP.stack((NP.arcsin(NP.clip(up, -1, 1)), NP.arctan2(east, north) % (2 * NP.pi)), axis=1)
    out = NP.degrees(out) if units == "degrees" else out
    return out.ravel() if shape1d else out


def altaz2hadec(altaz, latitude, units=None):
    shape1d = NP.asarray(altaz).ndim == 1
    altaz = NP.asarray(altaz, dtype=float).reshape(-1, 2)
    if units == "degrees":
        alt, az, lat = NP.radians(altaz[:, 0]), NP.radians(altaz[:, 1]), NP.radians(latitude)
    else:
        alt, az, lat = altaz[:, 0], altaz[:, 1], latitude
    e, n, u = NP.cos(alt) * NP.sin(az), NP.cos(alt) * NP.cos(az), NP.sin(alt)
    z = n * NP.cos(lat) + u * NP.sin(lat)
    x = -n * NP.sin(lat) + u * NP.cos(lat)
    out = NP.stack((NP.arctan2(-e, x) % (2 * NP.pi), NP.arcsin(NP.clip(z, -1, 1))), axis=1)
    out = NP.degrees(out) if units == "degrees" else out
    return out.ravel() if shape1d else out


def sphdist(lon1, lat1, lon2, lat2):
    lon1, lat1, lon2, lat2 = [NP.radians(NP.asarray(v, dtype=float)) for v in (lon1, lat1, lon2, lat2)]
    a = NP.sin(0.5 * (lat2 - lat1)) ** 2 + NP.cos(lat1) * NP.cos(lat2) * NP.sin(0.5 * (lon2 - lon1)) ** 2
    return NP.degrees(2 * NP.arcsin(NP.minimum(1.0, NP.sqrt(a))))


def spherematch(lon1, lat1, lon2, lat2, matchrad=None, nnearest=0, maxmatches=-1):
    """GEOM.spherematch [AU-memory] as used at interferometry.py:4549 / :6213 (one centre against many points,
    maxmatches=0): indices of the points within matchrad (great-circle, degrees) of the centre."""
    d = sphdist(NP.asarray(lon1, dtype=float).reshape(-1, 1), NP.asarray(lat1, dtype=float).reshape(-1, 1),
                NP.asarray(lon2, dtype=float).reshape(1, -1), NP.asarray(lat2, dtype=float).reshape(1, -1))
    m1, m2 = NP.where(d <= matchrad)
    return m1, m2, d[m1, m2]


def enu2xyz(enu, latitude, units=None):
    enu = NP.asarray(enu, dtype=float).reshape(-1, 3)
    lat = NP.radians(latitude) if units == "degrees" else latitude
    return NP.stack((-NP.sin(lat) * enu[:, 1] + NP.cos(lat) * enu[:, 2], enu[:, 0], NP.cos(lat) * enu[:, 1] + NP.sin(lat) * enu[:, 2]), axis=1)


def xyz2enu(xyz, latitude, units=None):
    xyz = NP.asarray(xyz, dtype=float).reshape(-1, 3)
    lat = NP.radians(latitude) if units == "degrees" else latitude
    return NP.stack((xyz[:, 1], -NP.sin(lat) * xyz[:, 0] + NP.cos(lat) * xyz[:, 2], NP.cos(lat) * xyz[:, 0] + NP.sin(lat) * xyz[:, 2]), axis=1)


def FT1D(inp, ax=-1, use_real=False, shift=False, inverse=False, verbose=True):
    out = NP.fft.ifft(inp, axis=ax) if inverse else NP.fft.fft(inp, axis=ax)
    return NP.fft.fftshift(out, axes=ax) if shift else out


def spectral_axis(length, delx=1.0, shift=False, use_real=False):
    ax = NP.fft.fftfreq(length, d=delx)
    return NP.fft.fftshift(ax) if shift else ax


def downsampler(inp, factor, axis=-1, verbose=True, method="interp", kind="linear", fill_value=NP.nan):
    inp = NP.asarray(inp)
    n = inp.shape[axis]
    if method in ("FFT", "fft"):        # [AU-memory]: Fourier resampling to round(n / factor) samples
        from scipy import signal
        return signal.resample(inp, int(NP.round(n / factor)), axis=axis)
    f = interpolate.interp1d(NP.arange(n), inp, kind=kind, axis=axis, bounds_error=False, fill_value=fill_value)
    return f(NP.arange(0, n, factor))


def windowing(N, shape="rect", pad_width=0, centering=True, area_normalize=False, peak=1.0, power_normalize=False):
    """DSP.windowing [AU-memory]: rect / 4-term Blackman-Harris / Blackman-Nuttall over n/(N-1); peak 1 unless normalised."""
    n = NP.arange(N)
    if shape.lower() == "rect":
        win = NP.ones(N)
    else:
        a = {"bhw": (0.35875, 0.48829, 0.14128, 0.01168), "bnw": (0.3635819, 0.4891775, 0.1365995, 0.0106411)}[shape.lower()]
        x = 2 * NP.pi * n / (N - 1)
        win = a[0] - a[1] * NP.cos(x) + a[2] * NP.cos(2 * x) - a[3] * NP.cos(3 * x)
    if area_normalize:
        return win / NP.sum(win)
    if power_normalize:
        return win / NP.sqrt(NP.sum(win ** 2))
    return win * peak / NP.amax(win)


def window_N2width(n_window=None, shape="rect", area_normalize=True, power_normalize=False, fftpow=1.0):
    """DSP.window_N2width [AU-memory]: width of the equivalent rectangular window as a fraction of the window length,
    sum(w / max w) / N on a long window (rect 1.0, bhw 0.35875, bnw 0.3635819)."""
    n = 1000000 if n_window is None else int(n_window)
    w = windowing(n, shape=shape) ** fftpow
    w = w / w.max()
    if power_normalize and not area_normalize:          # [AU-memory]: equivalent width of the POWER window
        return NP.sum(w ** 2) / n
    return NP.sum(w) / n


def window_fftpow(N_window, shape="rect", pad_width=0, pad_value=0.0, fftpow=1.0, centering=True, peak=None, area_normalize=False,
                  power_normalize=False, verbose=True):
    """DSP.window_fftpow [AU-memory] for fftpow = 1 only (the PRISim default): the plain window."""
    if fftpow != 1.0:
        raise NotImplementedError("window_fftpow stub: fftpow != 1 is not restated")
    return windowing(N_window, shape=shape, centering=centering, area_normalize=area_normalize, power_normalize=power_normalize,
                     peak=1.0 if peak is None else peak)


def find_1NN(ref, inp, distance_ULIM=NP.inf, remove_oob=True):
    """LKP.find_1NN [AU-memory]: nearest reference point of every input point; returns (input index, reference index,
    distance) of the inputs that have one within distance_ULIM."""
    ref = NP.asarray(ref, dtype=float).reshape(len(ref), -1); inp = NP.asarray(inp, dtype=float).reshape(len(inp), -1)
    d = NP.sqrt(((inp[:, None, :] - ref[None, :, :]) ** 2).sum(axis=2))
    j = NP.argmin(d, axis=1)
    dist = d[NP.arange(inp.shape[0]), j]
    ok = dist <= distance_ULIM
    return NP.arange(inp.shape[0])[ok], j[ok], dist[ok]


class SkyModel(object):
    """catalog.SkyModel stand-in: power-law ('func') spectra only."""

    def __init__(self, location, flux_scale, spindex, freq_ref, src_shape=None, epoch="J2000"):
        self.location = NP.asarray(location, dtype=float)
        self.flux_scale, self.spindex, self.freq_ref = map(lambda v: NP.asarray(v, dtype=float), (flux_scale, spindex, freq_ref))
        self.src_shape = src_shape
        self.epoch = epoch
        self.coords = "hadec"

    def generate_spectrum(self, ind=None, frequency=None, interp_method="pchip"):
        ind = NP.arange(self.location.shape[0]) if ind is None else NP.asarray(ind)
        f = NP.asarray(frequency, dtype=float).reshape(1, -1)
        return self.flux_scale[ind].reshape(-1, 1) * (f / self.freq_ref[ind].reshape(-1, 1)) ** self.spindex[ind].reshape(-1, 1)


def install_stubs():
    au = _mod("astroutils", __githash__="stub")
    au.geometry = _mod("astroutils.geometry", altaz2dircos=altaz2dircos, dircos2altaz=dircos2altaz, hadec2altaz=hadec2altaz,
                       altaz2hadec=altaz2hadec, sphdist=sphdist, xyz2enu=xyz2enu, enu2xyz=enu2xyz, spherematch=spherematch)
    au.DSP_modules = _mod("astroutils.DSP_modules", FT1D=FT1D, spectral_axis=spectral_axis, downsampler=downsampler,
                          windowing=windowing, window_N2width=window_N2width, window_fftpow=window_fftpow)
    au.catalog = _mod("astroutils.catalog", SkyModel=SkyModel)
    au.constants = _mod("astroutils.constants", Jy=1.0e-26, sday=0.99726958, rest_freq_HI=1420405751.77)
    for name in ("gridding_modules", "lookup_operations", "nonmathops", "mathops", "ephemeris_timing", "mpi_modules"):
        setattr(au, name, _mod("astroutils." + name))
    au.lookup_operations.find_1NN = find_1NN
    # NMO.find_all_occurrences_list1_in_list2 [AU-memory]: for every element of list1 the indices of list2 equal to it
    au.nonmathops.find_all_occurrences_list1_in_list2 = lambda l1, l2: [NP.where(NP.asarray(l2) == v)[0].tolist() for v in l1]

    class _Any(object):
        def __init__(self, *a, **k):
            pass

        def transform_to(self, *a, **k):
            return self

    ap = _mod("astropy")
    ap.io = _mod("astropy.io", fits=_mod("astropy.io.fits"), ascii=_mod("astropy.io.ascii"))
    ap.coordinates = _mod("astropy.coordinates", Galactic=_Any, SkyCoord=_Any, ICRS=_Any, FK5=_Any, AltAz=_Any, EarthLocation=_Any)
    ap.units = _mod("astropy.units", deg=1.0, m=1.0)
    ap.time = _mod("astropy.time", Time=_Any)
    _mod("h5py")
    _mod("progressbar")
    dv = _mod("distutils")
    dv.version = _mod("distutils.version", LooseVersion=_Any)
    _mod("prisim", __githash__="stub", __path__=[REF])


def load_reference(name):
    """exec a reference module from its source text with the Python-2 idioms patched in memory."""
    path = os.path.join(REF, name + ".py")
    src = open(path).read()
    src = re.sub(r"^(\s*)print '([^']*)'\s*$", r"\1print('\2')", src, flags=re.M)      # primary_beams.py:2027
    src = src.replace(".iteritems()", ".items()")                                       # interferometry.py:6405
    src = src.replace(".astype(NP.int)", ".astype(int)")                                # numpy >= 1.24 (interferometry.py:8238)
    src = src.replace("dtype=NP.float)", "dtype=float)")                                # numpy >= 1.24 (interferometry.py:957)
    src = src.replace("xy = zip(xloc, yloc)", "xy = list(zip(xloc, yloc))")             # interferometry.py:973 (zip is lazy in Py3)
    src = src.replace("NP.asarray(blgroups.keys(), dtype=self.labels.dtype)",           # interferometry.py:6858 (dict view)
                      "NP.asarray(list(blgroups.keys()), dtype=self.labels.dtype)")
    mod = types.ModuleType(name)
    mod.__file__ = path
    mod.__dict__["xrange"] = range
    sys.modules[name] = mod                       # implicit-relative imports (interferometry.py:27-28)
    exec(compile(src, path, "exec"), mod.__dict__)
    return mod


class TimeObj(object):
    def __init__(self, jd, lst_deg):
        self.jd = jd
        self._lst = lst_deg

    def sidereal_time(self, kind):
        return types.SimpleNamespace(deg=self._lst)


def main():
    if not os.path.isdir(REF):
        raise SystemExit("reference not present; golden vectors can only be regenerated in the build container")
    for alias, typ in (("int", int), ("float", float), ("bool", bool), ("complex", complex), ("float_", NP.float64)):
        if alias not in NP.__dict__:
            setattr(NP, alias, typ)             # NP.int / NP.float were removed from numpy
    install_stubs()
    DLY = load_reference("baseline_delay_horizon")
    PB = load_reference("primary_beams")
    RI = load_reference("interferometry")
    # delay_spectrum.py imports its siblings through the package and a few more third-party names at module level
    pr = sys.modules["prisim"]
    pr.primary_beams, pr.interferometry, pr.baseline_delay_horizon = PB, RI, DLY
    sys.modules.update({"prisim.primary_beams": PB, "prisim.interferometry": RI, "prisim.baseline_delay_horizon": DLY})

    class _Cosmo(object):
        H0, h = 67.7, 0.677

        def clone(self, **kw):
            return self
    sys.modules["astropy"].cosmology = _mod("astropy.cosmology", Planck15=_Cosmo(), WMAP9=_Cosmo())
    _mod("healpy")
    sys.modules["astroutils"].writer_module = _mod("astroutils.writer_module")
    DSM = load_reference("delay_spectrum")
    rng = NP.random.default_rng(20261017)
    lat = -30.7224

    # ---------------- beams (primary_beams.py executed as is) ----------------
    nsrc = 60
    altaz = NP.stack((rng.uniform(0.5, 89.5, nsrc), rng.uniform(0, 360, nsrc)), axis=1)
    altaz[0] = [90.0, 0.0]
    freqs_ghz = (150e6 + (NP.arange(16) - 8) * 2e6) / 1e9
    pc_altaz = NP.asarray([80.0, 30.0])
    beams = {}
    tel = {"hera": {"id": "hera", "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz"},
           "hirax_offzenith": {"id": "hirax", "orientation": NP.asarray([75.0, 120.0]), "ocoords": "altaz"},
           "mwa_dipole_gp": {"id": "mwa_dipole", "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3},
           "paper": {"id": "paper", "orientation": NP.asarray([0.0, 90.0]), "ocoords": "altaz"},
           "mwa_analytic": {"id": "mwa", "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3},
           "delta": {"shape": "delta"},
           "dish": {"shape": "dish", "size": 14.0, "ocoords": "altaz", "orientation": NP.asarray([90.0, 270.0])},
           "gaussian": {"shape": "gaussian", "size": 10.0, "ocoords": "altaz", "orientation": NP.asarray([90.0, 270.0])},
           "dipole_gp_mod": {"shape": "dipole", "size": 1.2, "ocoords": "dircos", "orientation": NP.asarray([[0.0, 1.0, 0.0]]),
                             "groundplane": 0.4, "ground_modify": {"scale": 0.8, "max": 1.5}}}
    for key, t in tel.items():
        kw = dict(skyunits="altaz", freq_scale="GHz")
        if key in ("dish", "gaussian"):
            kw["pointing_center"] = pc_altaz
        beams["pb_" + key] = PB.primary_beam_generator(altaz.copy(), freqs_ghz.copy(), dict(t), **kw)
    # dipole approximations through the wrapper
    beams["pb_paper_short"] = PB.primary_beam_generator(altaz.copy(), freqs_ghz.copy(), dict(tel["paper"]), skyunits="altaz", short_dipole_approx=True)
    beams["pb_paper_halfwave"] = PB.primary_beam_generator(altaz.copy(), freqs_ghz.copy(), dict(tel["paper"]), skyunits="altaz", half_wave_dipole_approx=True)
    # phased tile (array_field_pattern in the reference's float32), explicit delays then pointing-centre delays
    xl, yl = NP.meshgrid(1.1 * NP.linspace(-1.5, 1.5, 4), 1.1 * NP.linspace(1.5, -1.5, 4))
    element_locs = NP.hstack((xl.reshape(-1, 1), yl.reshape(-1, 1), NP.zeros((16, 1))))
    pcd = altaz2dircos(NP.asarray([[52.806, 101.31]]), "degrees")[0]
    delays = NP.dot(element_locs, pcd) / FCNST.c
    delays = NP.round((delays - delays.min()) / 435e-12) * 435e-12
    mwa_el = {"id": "mwa", "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3, "element_locs": element_locs}
    beams["pb_mwa_tile_delays"] = PB.primary_beam_generator(altaz.copy(), freqs_ghz.copy(), dict(mwa_el), skyunits="altaz",
                                                          pointing_info={"delays": delays.copy()})
    beams["pb_mwa_tile_pointing"] = PB.primary_beam_generator(altaz.copy(), freqs_ghz.copy(), dict(mwa_el), skyunits="altaz",
                                                            pointing_info={"pointing_center": NP.asarray([52.806, 101.31]), "pointing_coords": "altaz"})
    if not ONLY or "beams" in ONLY:
      NP.savez_compressed(os.path.join(OUT, "beams.npz"), altaz=altaz, freqs_ghz=freqs_ghz, pc_altaz=pc_altaz,
                        element_locs=element_locs, tile_delays=delays, **beams)

    # ---------------- geometric delays (baseline_delay_horizon.py executed as is) ----------------
    bl = rng.normal(0, 120.0, (9, 3)); bl[:, 2] *= 0.02
    hadec = NP.stack((rng.uniform(0, 360, 25), rng.uniform(-80, 40, 25)), axis=1)
    if not ONLY or "delays" in ONLY:
      NP.savez_compressed(os.path.join(OUT, "delays.npz"), bl=bl, altaz=altaz, hadec=hadec, latitude=lat,
                        tau_altaz=DLY.geometric_delay(bl, altaz, altaz=True, hadec=False),
                        tau_hadec=DLY.geometric_delay(bl, hadec, altaz=False, hadec=True, latitude=lat),
                        tau_dircos=DLY.geometric_delay(bl, altaz2dircos(altaz, "degrees"), altaz=False, hadec=False, dircos=True),
                        horizon=DLY.horizon_delay_limits(bl, altaz2dircos(NP.asarray([[90.0, 270.0]]), "degrees")))

    # ---------------- InterferometerArray.observe / generate_noise / add_noise / delay_transform ----------------
    def run_observe(tag, telescope, src_shape=None, roi_radius=None, pb_info=None, nbl=12, nchan=32, nsrc0=150, nsnap=3,
                    pointing_coords="hadec", roi_info_from_beam=False, gradient_mode=None):
        bl = rng.normal(0, 60.0, (nbl, 3)); bl[:, 2] *= 0.02
        bl[0] = [14.6, 0.0, 0.0]
        chans = 150e6 + (NP.arange(nchan) - nchan // 2) * 100e3
        labels = [("B{0}".format(i), "A{0}".format(i)) for i in range(nbl)]
        ra = rng.uniform(0, 360, nsrc0)
        dec = NP.degrees(NP.arcsin(rng.uniform(-1, 0.5, nsrc0)))
        flux = 10 ** rng.uniform(-1, 1.5, nsrc0)
        spindex = rng.normal(-0.83, 0.2, nsrc0)
        shp = None
        if src_shape is not None:
            fw = rng.uniform(0.05, src_shape, nsrc0)
            shp = NP.stack((fw, fw * rng.uniform(0.5, 1.0, nsrc0), NP.zeros(nsrc0)), axis=1)
        ia = RI.InterferometerArray(labels, bl, chans, telescope=dict(telescope), eff_Q=0.96, latitude=lat, longitude=21.4278,
                                    altitude=0.0, skycoords="hadec", A_eff=154.0 * 0.65, pointing_coords=pointing_coords,
                                    baseline_coords="localenu", freq_scale="Hz")
        Tsysinfo = {"Trx": 50.0, "Tant": {"T0": 200.0, "f0": 150e6, "spindex": -2.55}, "Tnet": None}
        bandpass = 1.0 + 0.1 * NP.cos(NP.arange(nchan) / 5.0)
        t_acc = [10.7, 10.7, 21.4][:nsnap]
        lsts = [0.0, 15.0, 40.0][:nsnap]
        pointing = NP.asarray([0.0, lat]) if pointing_coords == "hadec" else NP.asarray([60.0, 200.0])
        rec = {}
        for j in range(nsnap):
            hadec_j = NP.stack(((lsts[j] - ra) % 360.0, dec), axis=1)
            sky = SkyModel(hadec_j, flux, spindex, NP.full(nsrc0, 150e6), src_shape=shp)
            kw = {}
            if roi_info_from_beam:      # the run_prisim route: indices + externally computed beam table (:1959-1971)
                aa = hadec2altaz(hadec_j, lat, "degrees")
                ind = NP.where(aa[:, 0] >= 0.0)[0]
                pbt = PB.primary_beam_generator(aa[ind], chans / 1e9, dict(telescope), skyunits="altaz", freq_scale="GHz",
                                                pointing_center=hadec2altaz(pointing, lat, "degrees") if pointing_coords == "hadec" else pointing)
                kw["roi_info"] = {"ind": ind, "pbeam": pbt.astype(NP.float32)}
                rec["roi_ind_{0}".format(j)] = ind
                rec["roi_pbeam_{0}".format(j)] = pbt.astype(NP.float32)
            ia.observe(TimeObj(2451545.0 + j * 0.01, lsts[j]), Tsysinfo, bandpass, pointing, sky, t_acc[j], pb_info=pb_info,
                       roi_radius=roi_radius, gradient_mode=gradient_mode, **kw)
            rec["hadec_{0}".format(j)] = hadec_j
            rec["m2_{0}".format(j)] = NP.asarray(ia.obs_catalog_indices[-1]) if len(ia.obs_catalog_indices) > j else NP.zeros(0, dtype=int)
        if gradient_mode is not None:
            # visibility gradient w.r.t. baseline vectors (interferometry.py:6312-6343, :6384-6394) and the first-order
            # perturbed visibilities apply_gradients() builds from it (:6726-6819)
            pert = rng.normal(0, 0.02, (2, 3, nbl))
            # apply_gradients reads self.labels.size / self.lst.size (:6816), i.e. it expects the array attributes of an
            # object re-loaded from disk; observe() leaves them as lists
            keep = ia.labels, ia.lst
            ia.labels, ia.lst = NP.arange(nbl), NP.asarray(ia.lst)          # one label per baseline (a record array in the reference)
            delta = ia.apply_gradients(gradient_mode="baseline", perturbations={"baseline": pert.copy()})
            ia.labels, ia.lst = keep
            rec.update(gradient_baseline=ia.gradient["baseline"], perturbations=pert, delta_skyvis_freq=delta)
        NP.random.seed(1234 + len(tag))
        ia.generate_noise()
        ia.add_noise()
        window = nchan * windowing_bhw(nchan)
        ia.delay_transform(pad=1.0, freq_wts=window, verbose=False)
        rec.update(bl=bl, chans=chans, flux=flux, spindex=spindex, src_shape=(NP.zeros(0) if shp is None else shp), latitude=lat,
                   bandpass=bandpass, t_acc=NP.asarray(t_acc), lsts=NP.asarray(lsts), pointing=pointing, window=window,
                   noise_seed=1234 + len(tag), skyvis_freq=ia.skyvis_freq, vis_rms_freq=ia.vis_rms_freq, vis_noise_freq=ia.vis_noise_freq,
                   vis_freq=ia.vis_freq, Tsys=ia.Tsys, bp=ia.bp, lags=ia.lags, skyvis_lag=ia.skyvis_lag, vis_lag=ia.vis_lag,
                   lag_kernel=ia.lag_kernel, pointing_center=ia.pointing_center, n_acc=ia.n_acc, t_obs=ia.t_obs)
        # second pass of the transform without padding, and with pad=0.5 (non-integer decimation)
        ia.delay_transform(pad=0.0, freq_wts=window, verbose=False)
        rec["skyvis_lag_pad0"] = ia.skyvis_lag
        ia.delay_transform(pad=0.5, freq_wts=window, verbose=False)
        rec["skyvis_lag_pad05"] = ia.skyvis_lag
        if not ONLY or tag in ONLY:
            NP.savez_compressed(os.path.join(OUT, "observe_{0}.npz".format(tag)), **rec)
        if tag == "hera" and (not ONLY or "allruns_hera" in ONLY):
            # DelaySpectrum.delay_transform / delay_transform_allruns (delay_spectrum.py:1224-1342, :1475-1618) on the same object
            ds = DSM.DelaySpectrum(interferometer_array=ia)
            runs = NP.stack((ia.vis_freq, 0.5 * ia.skyvis_freq, 1j * ia.vis_noise_freq), axis=0).reshape(3, 1, nbl, nchan, nsnap)
            ar = {"runs": runs, "horizon_delay_limits": ds.horizon_delay_limits}
            r = ds.delay_transform_allruns(runs, pad=1.0, freq_wts=window, downsample=True, verbose=False)
            ar.update(vis_lag_pad1=r["vis_lag"], lag_kernel_pad1=r["lag_kernel"], lags_pad1=r["lags"])
            r = ds.delay_transform_allruns(ia.skyvis_freq, pad=0.5, freq_wts=None, downsample=False, verbose=False)
            ar.update(vis_lag_pad05_full=r["vis_lag"], lag_kernel_pad05_full=r["lag_kernel"], lags_pad05_full=r["lags"])
            wts2 = NP.outer(window, [1.0, 0.5, 0.25][:nsnap])                  # [nchan, nsnap] form of freq_wts
            r = ds.delay_transform_allruns(ia.skyvis_freq, pad=0.0, freq_wts=wts2, downsample=True, verbose=False)
            ar.update(vis_lag_pad0_w2=r["vis_lag"], wts2=wts2)
            r = ds.delay_transform(pad=1.0, freq_wts=window, downsample=True, action="return", verbose=False)
            ar.update(ds_skyvis_lag=r["skyvis_lag"], ds_vis_lag=r["vis_lag"], ds_lags=r["lags"], ds_lag_kernel=r["lag_kernel"])
            NP.savez_compressed(os.path.join(OUT, "allruns_hera.npz"), **ar)
        if tag == "hera" and (not ONLY or "multiwin_hera" in ONLY):
            # multi_window_delay_transform (interferometry.py:8141-8287): three sub-bands, Blackman-Harris, pad 1.0 and 0.0
            mw = {}
            bw_eff = NP.asarray([0.8e6, 1.0e6, 0.6e6]); fc = NP.asarray([chans[8], chans[16], chans[25]])
            for pad_mw, sfx in ((1.0, "pad1"), (0.0, "pad0")):
                res = ia.multi_window_delay_transform(bw_eff, freq_center=fc, shape="bhw", pad=pad_mw, verbose=False)
                for k in ("skyvis_lag", "vis_noise_lag", "lag_kernel", "lag_corr_length"):
                    mw["{0}_{1}".format(k, sfx)] = res[k]
            res = ia.multi_window_delay_transform(1.2e6, shape="rect", pad=1.0, verbose=False)      # defaults: centre channel
            mw.update(skyvis_lag_rect=res["skyvis_lag"], lag_corr_length_rect=res["lag_corr_length"], bw_eff=bw_eff, freq_center=fc,
                      vis_noise_freq=ia.vis_noise_freq)
            NP.savez_compressed(os.path.join(OUT, "multiwin_hera.npz"), **mw)
        if tag == "hera" and (not ONLY or "subband_hera" in ONLY):
            # DelaySpectrum.subband_delay_transform (delay_spectrum.py:1842-2248), 'sim' branch: three Blackman-Harris sub-bands
            # (one given out of order, one clipped by the band edge), pad 1.0; then one rectangular sub-band without padding
            ds2 = DSM.DelaySpectrum(interferometer_array=ia)
            sb = {}
            bw_sb = NP.asarray([0.9e6, 0.7e6, 1.1e6]); fc_sb = NP.asarray([chans[20] + 1.0e4, chans[9], chans[29] - 2.0e4])
            args = dict(freq_center={"cc": fc_sb.copy(), "sim": fc_sb.copy()}, shape={"cc": "bhw", "sim": "bhw"},
                        pad={"cc": 1.0, "sim": 1.0}, verbose=False)
            r = ds2.subband_delay_transform({"cc": bw_sb.copy(), "sim": bw_sb.copy()}, action="return_oversampled", **args)["sim"]
            for k in ("freq_wts", "lags", "skyvis_lag", "vis_lag", "vis_noise_lag", "lag_kernel", "lag_corr_length"):
                sb["bhw_" + k] = r[k]
            r = ds2.subband_delay_spectra_resampled["sim"]
            for k in ("lags", "skyvis_lag", "vis_lag", "vis_noise_lag", "lag_kernel", "lag_corr_length"):
                sb["bhw_rs_" + k] = r[k]
            r = ds2.subband_delay_transform({"cc": NP.asarray([1.3e6]), "sim": NP.asarray([1.3e6])},
                                            freq_center={"cc": NP.asarray([chans[15]]), "sim": NP.asarray([chans[15]])},
                                            pad={"cc": 0.0, "sim": 0.0}, action="return_resampled", verbose=False)["sim"]
            for k in ("lags", "skyvis_lag", "lag_kernel", "lag_corr_length"):
                sb["rect_rs_" + k] = r[k]
            sb.update(bw_eff=bw_sb, freq_center=fc_sb, rect_freq_wts=ds2.subband_delay_spectra["sim"]["freq_wts"],
                      rect_skyvis_lag=ds2.subband_delay_spectra["sim"]["skyvis_lag"],
                      skyvis_freq=ia.skyvis_freq, vis_freq=ia.vis_freq, vis_noise_freq=ia.vis_noise_freq, bp=ia.bp, chans=chans)
            NP.savez_compressed(os.path.join(OUT, "subband_hera.npz"), **sb)
        if tag == "hera" and (not ONLY or "rotate_hera" in ONLY):
            # rotate_visibilities = phase_centering + project_baselines (interferometry.py:7655-7995), twice:
            # to a fixed HA/Dec, then to an RA/Dec that differs per snapshot
            rot = {}
            ref1 = {"location": NP.asarray([[12.0, -24.0]]), "coords": "hadec"}
            ia.rotate_visibilities(ref1, do_delay_transform=False, verbose=False)
            rot.update(skyvis_rot1=ia.skyvis_freq, vis_rot1=ia.vis_freq, noise_rot1=ia.vis_noise_freq, pc_rot1=ia.phase_center,
                       proj_rot1=ia.projected_baselines)
            ref2 = {"location": NP.asarray([[350.0, -31.0], [5.0, -29.0], [30.0, -35.0]])[:nsnap], "coords": "radec"}
            ia.rotate_visibilities(ref2, do_delay_transform=False, verbose=False)
            rot.update(skyvis_rot2=ia.skyvis_freq, pc_rot2=ia.phase_center, proj_rot2=ia.projected_baselines,
                       ref1=ref1["location"], ref2=ref2["location"], pc_coords=NP.asarray(ia.phase_center_coords))
            NP.savez_compressed(os.path.join(OUT, "rotate_hera.npz"), **rot)

    hera = {"id": "hera", "shape": "dish", "size": 14.0, "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz", "groundplane": None}
    run_observe("hera", hera)
    run_observe("hera_taper", hera, src_shape=0.6)
    run_observe("hera_roi20", hera, roi_radius=20.0)
    run_observe("hera_roiinfo", hera, roi_info_from_beam=True, nsnap=2)
    run_observe("gaussian_altazpointing", {"shape": "gaussian", "size": 10.0, "ocoords": "altaz", "orientation": NP.asarray([90.0, 270.0]),
                                           "groundplane": None}, pointing_coords="altaz", nsnap=2)
    run_observe("mwa_dipole", {"id": "mwa_dipole", "shape": "dipole", "size": 0.74, "orientation": NP.asarray([1.0, 0.0, 0.0]),
                               "ocoords": "dircos", "groundplane": 0.3}, nsnap=2)
    # gradient_mode='baseline' only runs in the reference when the sky model has src_shape: the direction cosines it
    # multiplies by are assigned inside the taper branch (interferometry.py:6263) and are unbound otherwise (:6343)
    run_observe("hera_taper_gradient", hera, src_shape=0.6, gradient_mode="baseline", nsnap=2)
    # ---------------- thermalNoiseRMS (interferometry.py:89-230): broadcast shapes, both flux units; no random numbers ----------------
    if not ONLY or "thermal_rms" in ONLY:
        nb_, nc_, nt_ = 3, 4, 2
        Ts = 100.0 + NP.arange(nb_ * nc_ * nt_, dtype=float).reshape(nb_, nc_, nt_)
        tr = {"Tsys": Ts}
        tr["jy_full"] = RI.thermalNoiseRMS(12.5, 1e5, 10.7, Ts, nbl=nb_, nchan=nc_, ntimes=nt_, flux_unit="Jy", eff_Q=0.9)
        tr["k_full"] = RI.thermalNoiseRMS(12.5, 1e5, 10.7, Ts, nbl=nb_, nchan=nc_, ntimes=nt_, flux_unit="K", eff_Q=0.9)
        tr["jy_chan"] = RI.thermalNoiseRMS(NP.full((1, nc_, 1), 20.0), 1e5, 5.0, Ts[:1, :, :1], nbl=nb_, nchan=nc_, ntimes=nt_,
                                           eff_Q=NP.linspace(0.5, 1.0, nb_).reshape(nb_, 1, 1))
        tr["jy_scalar"] = RI.thermalNoiseRMS(100.0, 2e5, 1.0, 250.0)
        NP.savez_compressed(os.path.join(OUT, "thermal_rms.npz"), **tr)

    # ---------------- antenna layouts and baseline pairs (interferometry.py:857-989, :1184-1370): no random numbers ----------------
    if not ONLY or "layouts" in ONLY:
        lay = {}
        for tag, kw in (("side3", dict(n_side=3)), ("side11", dict(n_side=11)), ("side4_rot", dict(n_side=4, orientation=30.0, center=NP.asarray([[5.0, -3.0]])))):
            xy_ref, lab_ref = RI.hexagon_generator(14.6, **kw)
            lay["hex_" + tag] = xy_ref
            lay["hexlab_" + tag] = NP.asarray(list(lab_ref))
        ant3 = NP.hstack((lay["hex_side3"], 0.1 * NP.arange(19).reshape(-1, 1)))
        for tag, kw in (("plain", {}), ("auto", dict(auto=True)), ("conj", dict(conjugate=True))):
            b_ref, l_ref, i_ref = RI.baseline_generator(ant3, **kw)
            lay["bl_" + tag] = b_ref
            lay["bllab_" + tag] = NP.asarray([[x.decode() if isinstance(x, bytes) else str(x) for x in row] for row in l_ref.tolist()])
            lay["blid_" + tag] = NP.asarray([list(row) for row in i_ref.tolist()])
        NP.savez_compressed(os.path.join(OUT, "layouts.npz"), **lay)

    # ---------------- uniq_baselines (interferometry.py:1373-1463) on a redundant layout: no random numbers ----------------
    if not ONLY or "uniq_baselines" in ONLY:
        # 19-element hexagon, 14.6 m pitch (the reference's hexagon_generator needs Python 2's list-returning zip)
        xy = NP.asarray([[14.6 * (q + 0.5 * r), 14.6 * NP.sqrt(3.0) / 2 * r] for r in range(-2, 3) for q in range(-2, 3) if abs(q + r) <= 2])
        ants = NP.hstack((xy, NP.zeros((xy.shape[0], 1))))
        ants = NP.vstack((ants, [[100.3, -40.2, 1.5], [-77.7, 12.9, 0.0]]))                    # two outriggers
        ii, jj = NP.triu_indices(ants.shape[0], k=1)
        blu = ants[jj] - ants[ii]
        ub = {"bl": blu}
        for key, red in (("all", None), ("red", True), ("nonred", False)):
            sel, ind, cnt, occ = RI.uniq_baselines(blu, redundant=red)
            ub.update({"sel_" + key: sel, "ind_" + key: NP.asarray(ind), "cnt_" + key: NP.asarray(cnt),
                       "occ_len_" + key: NP.asarray([len(o) for o in occ]), "occ_flat_" + key: NP.concatenate([NP.asarray(o, dtype=int) for o in occ])})
        # group lookups (interferometry.py:2017-2165) on groups built as getBaselineInfo builds them (:1999-2007)
        dtl = [("A2", "U3"), ("A1", "U3")]
        labs = NP.asarray([(str(j), str(i)) for i, j in zip(ii, jj)], dtype=dtl)
        sel, ind, cnt, occ = RI.uniq_baselines(blu)
        grp, rev = {}, {}
        for gi, first in enumerate(ind):
            grp[tuple(labs[first])] = labs[NP.asarray(occ[gi])]
            for lbl in labs[NP.asarray(occ[gi])]:
                rev[tuple(lbl)] = NP.asarray([labs[first]], dtype=labs.dtype)
        query = [("1", "0"), ("0", "1"), ("20", "3"), ("7", "7"), ("5", "2"), ("2", "5"), ("19", "20")]
        keys, flip = RI.getBaselineGroupKeys(query, rev)
        members, flip2 = RI.getBaselinesInGroups(query, rev, grp)
        ub.update(query=NP.asarray(query), keys=NP.asarray([["", ""] if k is None else list(k) for k in keys]),
                  flipped=NP.asarray([-1 if f is None else int(f) for f in flip]),
                  member_counts=NP.asarray([-1 if m is None else m.size for m in members]),
                  member_first=NP.asarray([["", ""] if m is None else list(m[0].tolist()) for m in members]))
        NP.savez_compressed(os.path.join(OUT, "uniq_baselines.npz"), **ub)

    # ---------------- duplicate_measurements (interferometry.py:6823-6907): unique baselines -> redundant sets ----------------
    ant = NP.asarray([[0.0, 0.0, 0.0], [14.6, 0.0, 0.0], [29.2, 0.0, 0.0], [43.8, 0.0, 0.0], [0.0, 14.6, 0.0]])
    antl = NP.asarray(["0", "1", "2", "3", "4"])
    dt = [("A2", antl.dtype), ("A1", antl.dtype)]
    ulabels = NP.asarray([("1", "0"), ("2", "0"), ("3", "0"), ("4", "0")], dtype=dt)
    ubl = NP.asarray([[14.6, 0.0, 0.0], [29.2, 0.0, 0.0], [43.8, 0.0, 0.0], [0.0, 14.6, 0.0]])
    blgroups = {("1", "0"): NP.asarray([("1", "0"), ("2", "1"), ("3", "2")], dtype=dt),      # key listed in its own group
                ("2", "0"): NP.asarray([("3", "1")], dtype=dt),                              # key missing: prepended (:6866-6870)
                ("3", "0"): NP.asarray([("3", "0")], dtype=dt)}                              # singleton; ("4","0") has no group at all
    nchan, nsrc0 = 16, 80
    chans = 150e6 + (NP.arange(nchan) - nchan // 2) * 100e3
    ia = RI.InterferometerArray(ulabels, ubl, chans, telescope=dict(hera), eff_Q=0.96, latitude=lat, longitude=21.4278, altitude=0.0,
                                skycoords="hadec", A_eff=154.0 * 0.65, pointing_coords="hadec", baseline_coords="localenu", freq_scale="Hz",
                                layout={"positions": ant, "labels": antl, "ids": NP.arange(5), "coords": "ENU"},
                                blgroupinfo={"groups": blgroups, "reversemap": None})
    ra = rng.uniform(0, 360, nsrc0); dec = NP.degrees(NP.arcsin(rng.uniform(-1, 0.5, nsrc0)))
    flux = 10 ** rng.uniform(-1, 1.5, nsrc0); spindex = rng.normal(-0.83, 0.2, nsrc0)
    rec = dict(bl=ubl, chans=chans, flux=flux, spindex=spindex, latitude=lat, lsts=NP.asarray([0.0, 20.0]),
               ulabels=NP.asarray([list(l) for l in ulabels.tolist()]),
               group_keys=NP.asarray([list(k) for k in blgroups]),
               group_0=NP.asarray([list(l) for l in blgroups[("1", "0")].tolist()]),
               group_1=NP.asarray([list(l) for l in blgroups[("2", "0")].tolist()]),
               group_2=NP.asarray([list(l) for l in blgroups[("3", "0")].tolist()]))
    for j, lst_j in enumerate((0.0, 20.0)):
        hadec_j = NP.stack(((lst_j - ra) % 360.0, dec), axis=1)
        rec["hadec_{0}".format(j)] = hadec_j
        ia.observe(TimeObj(2451545.0 + j * 0.01, lst_j), {"Trx": 50.0, "Tant": {"T0": 200.0, "f0": 150e6, "spindex": -2.55}, "Tnet": None},
                   NP.ones(nchan), NP.asarray([0.0, lat]), SkyModel(hadec_j, flux, spindex, NP.full(nsrc0, 150e6)), 10.7)
    rec["skyvis_unique"] = ia.skyvis_freq.copy()
    # duplicate_measurements repeats projected_baselines too (:6894), which only exist after project_baselines (:7916-7995)
    ia.project_baselines(ref_point={"location": NP.asarray([[0.0, lat]]), "coords": "hadec"})
    rec["projected_unique"] = ia.projected_baselines.copy()
    NP.random.seed(99)
    ia.duplicate_measurements()
    rec.update(labels_out=NP.asarray([list(l) for l in ia.labels.tolist()]), baselines_out=ia.baselines, skyvis_out=ia.skyvis_freq,
               baseline_lengths_out=ia.baseline_lengths, projected_out=ia.projected_baselines, Tsys_out=ia.Tsys, vis_rms_out=ia.vis_rms_freq, bp_out=ia.bp,
               vis_noise_shape=NP.asarray(ia.vis_noise_freq.shape), vis_minus_noise=ia.vis_freq - ia.vis_noise_freq)
    if not ONLY or "duplicate" in ONLY:
        NP.savez_compressed(os.path.join(OUT, "duplicate.npz"), **rec)
    # ---------------- ROI_parameters.append_settings (interferometry.py:4221-4617): per-snapshot ROI indices + beam table ----------------
    nsrc0 = 300
    hadec0 = NP.stack((rng.uniform(0, 360, nsrc0), NP.degrees(NP.arcsin(rng.uniform(-1, 0.6, nsrc0)))), axis=1)
    roi_sky = SkyModel(hadec0, NP.ones(nsrc0), NP.zeros(nsrc0), NP.full(nsrc0, 150e6))
    roi_freq = 150e6 + (NP.arange(24) - 12) * 250e3
    roirec = dict(hadec=hadec0, freq=roi_freq, latitude=lat)
    tel_roi = dict(hera); tel_roi.update(latitude=lat, longitude=21.4278, altitude=0.0)
    cases = [("zenith_achromatic", {"radius": None, "center": None, "center_coords": None}),
             ("zenith_r30_chromatic", {"radius": 30.0, "center": None, "center_coords": None, "pbeam_chromaticity": True}),
             ("offzenith_altaz_reffreq", {"radius": 25.0, "center": NP.asarray([60.0, 140.0]), "center_coords": "altaz", "pbeam_reffreq": 151.3e6}),
             ("offzenith_hadec", {"radius": 40.0, "center": NP.asarray([20.0, -10.0]), "center_coords": "hadec"}),
             ("given_ind", {"ind": NP.asarray([3, 17, 44, 120, 250]), "radius": 90.0}),
             ("given_ind_pbeam", {"ind": NP.asarray([5, 6, 7]), "pbeam": rng.uniform(0, 1, (3, 24))})]
    roi = RI.ROI_parameters()
    for name, ri in cases:
        roi.append_settings(roi_sky, roi_freq, pinfo={"pointing_center": NP.asarray([[90.0, 270.0]]), "pointing_coords": "altaz"},
                            lst=10.0, time_jd=2451545.0, roi_info=dict(ri), telescope=tel_roi, freq_scale="Hz")
        roirec["ind_" + name] = NP.asarray(roi.info["ind"][-1]); roirec["pbeam_" + name] = NP.asarray(roi.info["pbeam"][-1])
        if "radius" in ri and len(roi.info["radius"]):
            roirec["radius_" + name] = NP.asarray(roi.info["radius"][-1], dtype=float)
        if "pbeam" in ri:
            roirec["pbeam_in_" + name] = ri["pbeam"]
    roi.append_settings(None, roi_freq, pinfo=None, roi_info=None, telescope=tel_roi, freq_scale="Hz")
    roirec["n_entries"] = len(roi.info["ind"]); roirec["center_coords"] = str(roi.info["center_coords"])
    roirec["centers"] = NP.concatenate([NP.asarray(c, dtype=float).reshape(1, 2) for c in roi.info["center"]], axis=0)
    if not ONLY or "roi_parameters" in ONLY:
        NP.savez_compressed(os.path.join(OUT, "roi_parameters.npz"), **roirec)
    # ---------------- Tsys / bandpass bookkeeping forms of observe() (interferometry.py:5993-6086) ----------------
    nbl, nchan, nsrc0 = 5, 16, 60
    blk = rng.normal(0, 40.0, (nbl, 3)); blk[:, 2] *= 0.02
    chans = 150e6 + (NP.arange(nchan) - nchan // 2) * 100e3
    ia = RI.InterferometerArray([("B{0}".format(i), "A{0}".format(i)) for i in range(nbl)], blk, chans, telescope=dict(hera), eff_Q=0.9,
                                latitude=lat, skycoords="hadec", A_eff=NP.linspace(50.0, 90.0, nbl), pointing_coords="hadec",
                                baseline_coords="localenu", freq_scale="Hz")
    hadec_k = NP.stack((rng.uniform(0, 360, nsrc0), NP.degrees(NP.arcsin(rng.uniform(-1, 0.5, nsrc0)))), axis=1)
    flux_k = 10 ** rng.uniform(-1, 1, nsrc0); sp_k = rng.normal(-0.8, 0.2, nsrc0)
    sky_k = SkyModel(hadec_k, flux_k, sp_k, NP.full(nsrc0, 150e6))
    bp1 = 1.0 + 0.2 * NP.sin(NP.arange(nchan) / 3.0)
    bp2 = rng.uniform(0.5, 1.5, (nbl, nchan))
    bp3 = rng.uniform(0.5, 1.5, (nbl, nchan, 1))
    forms = [(bp1, {"Tnet": 150.0}, None),
             (bp2, {"Tnet": NP.linspace(100.0, 200.0, nbl)}, None),
             (bp3, {"Trx": 40.0, "Tant": {"T0": 180.0, "f0": 150e6, "spindex": -2.5}, "Tnet": None}, 1.0 + 0.05 * NP.arange(nchan)),
             (bp1, {"Tnet": NP.linspace(90.0, 120.0, nchan)}, NP.linspace(0.8, 1.2, nbl)),
             (bp1, {"Trx": 40.0, "Tant": {"T0": 180.0, "f0": 150e6, "spindex": -2.5}}, rng.uniform(0.9, 1.1, (nbl, nchan)))]
    for j, (bpj, tsj, bcj) in enumerate(forms):
        ia.observe(TimeObj(2451545.0 + j * 0.01, 5.0 * j), tsj, bpj, NP.asarray([0.0, lat]), sky_k, 10.0 + j, bpcorrect=bcj)
    NP.random.seed(3)
    ia.generate_noise()
    ia.add_noise()
    wk = nchan * windowing_bhw(nchan)
    ia.delay_transform(pad=1.0, freq_wts=wk, verbose=False)
    if not ONLY or "bookkeeping" in ONLY:
        NP.savez_compressed(os.path.join(OUT, "bookkeeping.npz"), bl=blk, chans=chans, hadec=hadec_k, flux=flux_k, spindex=sp_k, latitude=lat,
                            bp1=bp1, bp2=bp2, bp3=bp3, bc2=forms[2][2], bc3=forms[3][2], bc4=forms[4][2], A_eff=NP.linspace(50.0, 90.0, nbl),
                            Tsys=ia.Tsys, bp=ia.bp, bp_wts=ia.bp_wts, vis_rms_freq=ia.vis_rms_freq, skyvis_freq=ia.skyvis_freq,
                            skyvis_lag=ia.skyvis_lag, lag_kernel=ia.lag_kernel, t_acc=NP.asarray(ia.t_acc), window=wk)
    print("golden vectors written to", OUT)


def windowing_bhw(N):
    n = NP.arange(N)
    x = 2 * NP.pi * n / (N - 1)
    w = 0.35875 - 0.48829 * NP.cos(x) + 0.14128 * NP.cos(2 * x) - 0.01168 * NP.cos(3 * x)
    return w / w.sum()


if __name__ == "__main__":
    main()
