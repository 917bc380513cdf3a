"""TEST INFRASTRUCTURE ONLY -- float64 numpy restatement of the PRISim visibility hot path.

This module is the parity oracle and the timed CPU baseline.  It is never imported by the
product package ``prisim_b200`` (see ``oracle/__init__.py``).

Every function cites the reference lines (relative to ``/root/reference/``) it follows.  The
reference (PRISim v2.2.1) is Python-2 / numpy code that cannot be imported unmodified here
(missing ``astroutils``, ``astropy``, ``h5py``, ...; one Python-2 ``print`` statement).

Pinning status
--------------
* The reference ships no tests, golden vectors or fixtures.
* ``tests/golden/make_golden.py`` executes the reference's OWN source files
  (``prisim/primary_beams.py``, ``prisim/baseline_delay_horizon.py`` and the
  ``InterferometerArray.observe / generate_noise / delay_transform`` methods of
  ``prisim/interferometry.py``) in this container under Python 3, with stub modules standing
  in for the absent third-party packages, and commits the outputs as
  ``tests/golden/*.npz``.  ``tests/test_oracle_golden.py`` checks this oracle against them.
* What stays unpinned: the semantics of the un-vendored ``astroutils`` helpers (marked
  [AU-memory] below -- restated from recollection of that package's public behaviour; the
  golden generator uses the same restatement for its stubs) and astropy's apparent-place
  pipeline (outside the parity boundary: the device contract starts at HA/Dec or Alt/Az).

Shapes follow the reference: sources x channels for beams/spectra, nbl x nchan (x nsnap) for
visibilities.  Everything is float64 / complex128.
"""
from __future__ import annotations

import numpy as NP
import scipy.constants as FCNST
import scipy.special as SPS
from scipy import interpolate

Jy = 1.0e-26  # CNST.Jy [AU-memory]


# --------------------------------------------------------------------------------------------
# [AU-memory] astroutils.geometry
# --------------------------------------------------------------------------------------------
def altaz2dircos(altaz, units="degrees"):
    """GEOM.altaz2dircos [AU-memory]: (l,m,n) = (cos alt sin az, cos alt cos az, sin alt); az is
    measured from North through East, axes are East, North, Up.  Call sites:
    baseline_delay_horizon.py:218, interferometry.py:6164."""
    altaz = NP.asarray(altaz, dtype=NP.float64).reshape(-1, 2)
    if units == "degrees":
        altaz = NP.radians(altaz)
    alt, az = altaz[:, 0], altaz[:, 1]
    return NP.stack((NP.cos(alt) * NP.sin(az), NP.cos(alt) * NP.cos(az), NP.sin(alt)), axis=1)


def dircos2altaz(dircos, units="degrees"):
    """GEOM.dircos2altaz [AU-memory]: inverse of altaz2dircos (primary_beams.py:598, :952)."""
    dircos = NP.asarray(dircos, dtype=NP.float64).reshape(-1, 3)
    alt = NP.arcsin(NP.clip(dircos[:, 2], -1.0, 1.0))
    az = NP.arctan2(dircos[:, 0], dircos[:, 1]) % (2 * NP.pi)
    out = NP.stack((alt, az), axis=1)
    return NP.degrees(out) if units == "degrees" else out


def hadec2altaz(hadec, latitude, units="degrees"):
    """GEOM.hadec2altaz [AU-memory]: spherical triangle at geodetic latitude; hour angle is
    positive to the West; azimuth North through East in [0, 360).  Call sites:
    interferometry.py:6157, :6177; baseline_delay_horizon.py:220."""
    hadec = NP.asarray(hadec, dtype=NP.float64).reshape(-1, 2)
    if units == "degrees":
        ha, dec, lat = NP.radians(hadec[:, 0]), NP.radians(hadec[:, 1]), NP.radians(latitude)
    else:
        ha, dec, lat = hadec[:, 0], hadec[:, 1], latitude
    north = NP.sin(dec) * NP.cos(lat) - NP.cos(dec) * NP.cos(ha) * NP.sin(lat)
    east = -NP.cos(dec) * NP.sin(ha)
    up = NP.sin(dec) * NP.sin(lat) + NP.cos(dec) * NP.cos(ha) * NP.cos(lat)
    alt = NP.arcsin(NP.clip(up, -1.0, 1.0))
    az = NP.arctan2(east, north) % (2 * NP.pi)
    out = NP.stack((alt, az), axis=1)
    return NP.degrees(out) if units == "degrees" else out


def altaz2hadec(altaz, latitude, units="degrees"):
    """GEOM.altaz2hadec [AU-memory]: inverse of hadec2altaz (interferometry.py:6122)."""
    altaz = NP.asarray(altaz, dtype=NP.float64).reshape(-1, 2)
    if units == "degrees":
        alt, az, lat = NP.radians(altaz[:, 0]), NP.radians(altaz[:, 1]), NP.radians(latitude)
    else:
        alt, az, lat = altaz[:, 0], altaz[:, 1], latitude
    # ENU unit vector -> equatorial (HA, Dec)
    e, n, u = NP.cos(alt) * NP.sin(az), NP.cos(alt) * NP.cos(az), NP.sin(alt)
    z = n * NP.cos(lat) + u * NP.sin(lat)           # towards the celestial pole
    x = -n * NP.sin(lat) + u * NP.cos(lat)          # towards HA=0 on the equator
    y = -e                                           # towards HA=+6h (West)
    dec = NP.arcsin(NP.clip(z, -1.0, 1.0))
    ha = NP.arctan2(y, x) % (2 * NP.pi)
    out = NP.stack((ha, dec), axis=1)
    return NP.degrees(out) if units == "degrees" else out


def sphdist(lon1, lat1, lon2, lat2):
    """GEOM.sphdist [AU-memory]: great-circle separation, all angles in degrees
    (primary_beams.py:605, :712)."""
    lon1, lat1, lon2, lat2 = [NP.radians(NP.asarray(x, dtype=NP.float64)) for x in (lon1, lat1, lon2, lat2)]
    a = NP.sin(0.5 * (lat2 - lat1)) ** 2 + NP.cos(lat1) * NP.cos(lat2) * NP.sin(0.5 * (lon2 - lon1)) ** 2
    return NP.degrees(2.0 * NP.arcsin(NP.minimum(1.0, NP.sqrt(a))))


def xyz2enu(xyz, latitude, units="degrees"):
    """GEOM.xyz2enu [AU-memory]: equatorial (X towards HA=0/Dec=0, Y towards HA=-6h i.e. East,
    Z towards the pole) -> local East, North, Up at the latitude (interferometry.py:6153)."""
    xyz = NP.asarray(xyz, dtype=NP.float64).reshape(-1, 3)
    lat = NP.radians(latitude) if units == "degrees" else latitude
    e = xyz[:, 1]
    n = -NP.sin(lat) * xyz[:, 0] + NP.cos(lat) * xyz[:, 2]
    u = NP.cos(lat) * xyz[:, 0] + NP.sin(lat) * xyz[:, 2]
    return NP.stack((e, n, u), axis=1)


# --------------------------------------------------------------------------------------------
# baseline_delay_horizon.py
# --------------------------------------------------------------------------------------------
def geometric_delay(baselines, skypos, altaz=False, dircos=False, hadec=True, units="mks", latitude=None):
    """DLY.geometric_delay, baseline_delay_horizon.py:133-241.  Returns [nsrc, nbl] seconds:
    ``NP.dot(dc, baselines.T)/c`` (:240)."""
    baselines = NP.asarray(baselines, dtype=NP.float64)
    if baselines.ndim == 1:
        baselines = baselines.reshape(1, -1)                       # :195-196
    if baselines.shape[1] > 3:
        baselines = baselines[:, :3]                               # :202-203
    skypos = NP.asarray(skypos, dtype=NP.float64)
    if altaz or hadec:                                             # :205-220
        if skypos.ndim < 2:
            skypos = skypos.reshape(1, -1)
        if altaz:
            dc = altaz2dircos(skypos, "degrees")
        else:
            dc = altaz2dircos(hadec2altaz(skypos, latitude, "degrees"), "degrees")
    else:                                                          # :221-233
        dc = skypos.reshape(1, -1) if skypos.ndim < 2 else skypos
    c = FCNST.c if units == "mks" else FCNST.c * 1e2               # :236-237
    return NP.dot(dc, baselines.T) / c                             # :240


def horizon_delay_limits(baselines, refdir_dircos):
    """baseline_delay_horizon.py:91-94, :128: [-|b|/c - b.s_pc/c, +|b|/c - b.s_pc/c] per baseline."""
    baselines = NP.asarray(baselines, dtype=NP.float64).reshape(-1, 3)
    blen = NP.sqrt(NP.sum(baselines ** 2, axis=1))
    off = NP.dot(baselines, NP.asarray(refdir_dircos, dtype=NP.float64).reshape(3)) / FCNST.c
    return NP.stack((-blen / FCNST.c - off, blen / FCNST.c - off), axis=1)


# --------------------------------------------------------------------------------------------
# primary_beams.py pattern functions
# --------------------------------------------------------------------------------------------
def _angle_from_pointing(skypos, skyunits, pointing_center, pointing_coords):
    """Common front end of airy_disk_pattern / gaussian_beam (primary_beams.py:575-607 and
    :682-714): angular offset x [rad] from the pointing centre and the beyond-horizon mask."""
    skypos = NP.asarray(skypos, dtype=NP.float64)
    if pointing_center is None:                                    # :575-584
        if skyunits == "altaz":
            x = NP.radians(90.0 - skypos[:, 0])
        elif skyunits == "dircos":
            x = NP.arcsin(NP.sqrt(skypos[:, 0] ** 2 + skypos[:, 1] ** 2))
        else:
            raise ValueError("skyunits must be altaz or dircos")
        zero_ind = x >= NP.pi / 2
    else:                                                          # :585-607
        if pointing_coords is None:
            pointing_coords = skyunits
        pc_altaz = NP.asarray(pointing_center, dtype=NP.float64).reshape(1, -1)
        if pointing_coords == "dircos":
            pc_altaz = dircos2altaz(pc_altaz, units="degrees")
        skypos_altaz = NP.copy(skypos)
        if skyunits == "dircos":
            skypos_altaz = dircos2altaz(skypos, units="degrees")
        x = sphdist(skypos_altaz[:, 1], skypos_altaz[:, 0], pc_altaz[0, 1], pc_altaz[0, 0])
        x = NP.radians(x)
        zero_ind = NP.logical_or(x >= NP.pi / 2, skypos_altaz[:, 0] <= 0.0)
    return x, zero_ind


def airy_disk_pattern(diameter, skypos, frequency, skyunits="altaz", peak=1.0, pointing_center=None,
                      pointing_coords=None, small_angle_tol=1e-10, power=True):
    """primary_beams.py:517-625 (frequency in Hz)."""
    frequency = NP.asarray(frequency, dtype=NP.float64).ravel()
    x, zero_ind = _angle_from_pointing(skypos, skyunits, pointing_center, pointing_coords)
    k = (2 * NP.pi * frequency / FCNST.c).reshape(1, -1)           # :609-610
    x = NP.where(x < small_angle_tol, small_angle_tol, x).reshape(-1, 1)   # :611-613
    arg = k * 0.5 * diameter * NP.sin(x)
    pattern = 2 * SPS.j1(arg) / arg                                # :614
    pattern[zero_ind, :] = 0.0                                     # :616
    arg0 = k * 0.5 * diameter * NP.sin(small_angle_tol)
    maxval = 2 * SPS.j1(arg0) / arg0                               # :618
    if power:                                                      # :619-621
        pattern = NP.abs(pattern) ** 2
        maxval = maxval ** 2
    pattern *= peak / maxval                                       # :623
    return pattern


def gaussian_beam(diameter, skypos, frequency, skyunits="altaz", pointing_center=None,
                  pointing_coords=None, power=True):
    """primary_beams.py:629-730."""
    frequency = NP.asarray(frequency, dtype=NP.float64).ravel()
    x, zero_ind = _angle_from_pointing(skypos, skyunits, pointing_center, pointing_coords)
    x = x.reshape(-1, 1)                                           # :716
    sigma_aprtr = diameter / (2.0 * NP.sqrt(2.0 * NP.log(2.0))) / (FCNST.c / frequency)   # :717
    sigma_dircos = (1.0 / (2 * NP.pi * sigma_aprtr)).reshape(1, -1)                      # :721-722
    pattern = NP.exp(-0.5 * (NP.sin(x) / sigma_dircos) ** 2)       # :723-724
    pattern[zero_ind, :] = 0.0                                     # :725
    if power:
        pattern = NP.abs(pattern) ** 2                             # :727-728
    return pattern


def _skypos_to_dircos(skypos, skycoords):
    skypos = NP.asarray(skypos, dtype=NP.float64)
    if skycoords == "altaz":
        return altaz2dircos(skypos.reshape(-1, 2), units="degrees")
    return skypos.reshape(-1, 3)


def dipole_field_pattern(length, skypos, dipole_coords=None, dipole_orientation=None, skycoords=None,
                         wavelength=1.0, short_dipole_approx=False, half_wave_dipole_approx=True,
                         power=False):
    """primary_beams.py:975-1235 (field pattern, [nsrc, nchan])."""
    wavelength = NP.asarray(wavelength, dtype=NP.float64).reshape(-1)
    if dipole_coords is None:
        dipole_coords = skycoords
    if skycoords is None:
        skycoords = dipole_coords
    skypos_dircos = _skypos_to_dircos(skypos, skycoords)           # :1112-1155
    if dipole_orientation is None:
        orient = NP.asarray([1.0, 0.0, 0.0]).reshape(1, -1)        # :1201 (default east)
    elif dipole_coords == "altaz":
        orient = altaz2dircos(NP.asarray(dipole_orientation, dtype=NP.float64).reshape(1, 2), units="degrees")   # :1178
    else:
        orient = NP.asarray(dipole_orientation, dtype=NP.float64).reshape(1, 3)
    k = 2 * NP.pi / wavelength.reshape(1, -1)                      # :1205
    h = 0.5 * length
    dot_product = NP.dot(orient, skypos_dircos.T).reshape(-1, 1)   # :1207
    angles = NP.arccos(dot_product)                                # :1208
    eps = 1.0e-10
    zero_angles_ind = NP.abs(NP.abs(dot_product) - 1.0) < eps      # :1211
    max_pattern = 1.0
    if short_dipole_approx:                                        # :1216-1218
        field_pattern = NP.repeat(NP.sin(angles).reshape(-1, 1), wavelength.size, axis=1)
    else:
        if half_wave_dipole_approx:                                # :1220-1222
            field_pattern = NP.cos(0.5 * NP.pi * NP.cos(angles)) / NP.sin(angles)
            field_pattern = NP.repeat(field_pattern.reshape(-1, 1), wavelength.size, axis=1)
        else:                                                      # :1224-1225
            max_pattern = 1.0 - NP.cos(k * h)
            with NP.errstate(divide="ignore", invalid="ignore"):
                field_pattern = (NP.cos(k * h * NP.cos(angles)) - NP.cos(k * h)) / NP.sin(angles)
        if NP.sum(zero_angles_ind) > 0:                            # :1227-1228
            field_pattern[zero_angles_ind.ravel(), :] = (
                k * h * NP.sin(k * h * NP.cos(angles[zero_angles_ind]).reshape(-1, 1))
                * NP.tan(angles[zero_angles_ind]).reshape(-1, 1))
    if power:
        return NP.abs(field_pattern / max_pattern) ** 2
    return field_pattern / max_pattern


def ground_plane_field_pattern(height, skypos, skycoords=None, wavelength=1.0, angle_units=None,
                               modifier=None, power=True):
    """primary_beams.py:812-971."""
    wavelength = NP.asarray(wavelength, dtype=NP.float64).reshape(-1)
    skypos_dircos = _skypos_to_dircos(skypos, skycoords)
    k = 2 * NP.pi / wavelength                                     # :950
    skypos_altaz = dircos2altaz(skypos_dircos, units="radians")    # :952
    ground_pattern = 2 * NP.sin(k.reshape(1, -1) * height * NP.sin(skypos_altaz[:, 0].reshape(-1, 1)))   # :953
    if isinstance(modifier, dict):                                 # :955-963
        val = 1.0 / NP.sqrt(NP.abs(skypos_dircos[:, 2]))
        if "scale" in modifier:
            val = val * modifier["scale"]
        if "max" in modifier:
            val = NP.clip(val, 0.0, modifier["max"])
        ground_pattern = ground_pattern * val[:, NP.newaxis]
    max_pattern = 2 * NP.sin(k.reshape(1, -1) * height * NP.sin(NP.pi / 2))   # :965
    ground_pattern = ground_pattern / max_pattern                  # :966
    if power:
        return NP.abs(ground_pattern) ** 2
    return ground_pattern


def isotropic_radiators_array_field_pattern(nax1, nax2, sep1, sep2=None, skypos=None, wavelength=1.0,
                                            east2ax1=None, skycoords="altaz", pointing_center=None,
                                            power=True):
    """primary_beams.py:1239-1478 for the call made by the 'mwa' preset (:282-285): skycoords
    'altaz' or 'dircos', east2ax1 a number.  Mirrors the reference's use of nax1 on both axes
    (:1471)."""
    wavelength = NP.asarray(wavelength, dtype=NP.float64).reshape(-1)
    if sep2 is None:
        sep2 = sep1
    skypos = NP.asarray(skypos, dtype=NP.float64)
    if east2ax1 is None:
        east2ax1 = 0.0
    if skycoords == "altaz":                                       # :1434-1442
        if pointing_center is None:
            pointing_center = NP.asarray([90.0, 0.0])
        pointing_center = NP.asarray(pointing_center, dtype=NP.float64).ravel()
        rot = altaz2dircos(NP.hstack((skypos[:, 0].reshape(-1, 1), (skypos[:, 1] + east2ax1).reshape(-1, 1))), units="degrees")
        pc_rot = altaz2dircos([pointing_center[0], pointing_center[1] + east2ax1], units="degrees")
    else:                                                          # :1443-1449
        if pointing_center is None:
            pointing_center = NP.asarray([0.0, 0.0, 1.0])
        pointing_center = NP.asarray(pointing_center, dtype=NP.float64).reshape(1, 3)
        angle = NP.radians(east2ax1)
        rotation_matrix = NP.asarray([[NP.cos(angle), NP.sin(angle), 0.0],
                                      [-NP.sin(angle), NP.cos(angle), 0.0],
                                      [0.0, 0.0, 1.0]])
        rot = NP.dot(skypos, rotation_matrix.T)
        pc_rot = NP.dot(pointing_center, rotation_matrix.T)
    rel = rot - pc_rot.reshape(1, -1)                              # :1451
    phi = 2 * NP.pi * sep1 * rel[:, 0].reshape(-1, 1) / wavelength.reshape(1, -1)   # :1460
    psi = 2 * NP.pi * sep2 * rel[:, 1].reshape(-1, 1) / wavelength.reshape(1, -1)   # :1461
    eps = 1.0e-10
    zero_phi = NP.abs(phi) < eps
    zero_psi = NP.abs(psi) < eps
    with NP.errstate(divide="ignore", invalid="ignore"):
        term1 = NP.sin(0.5 * nax1 * phi) / NP.sin(0.5 * phi) / nax1            # :1467
        term2 = NP.sin(0.5 * nax1 * psi) / NP.sin(0.5 * psi) / nax1            # :1471 (nax1, sic)
    term1[zero_phi] = NP.cos(0.5 * nax1 * phi[zero_phi]) / NP.cos(0.5 * phi[zero_phi])   # :1468-1469
    term2[zero_psi] = NP.cos(0.5 * nax1 * psi[zero_psi]) / NP.cos(0.5 * psi[zero_psi])   # :1472-1473
    pb = term1 * term2
    if power:
        pb = NP.abs(pb) ** 2
    return pb


def array_field_pattern(antpos, skypos, skycoords=None, pointing_info=None, wavelength=1.0, power=True,
                        reference_float32=False, randn=None):
    """primary_beams.py:1482-1754: (1/N) sum_e g_e exp(2 pi i f (-r_e.s/c + d_e)), shape
    [nsrc, nchan, nrand].  The reference evaluates this in float32/complex64 (:1593, :1718,
    :1730-1746); ``reference_float32=True`` reproduces that, the default float64 evaluation is
    the parity target.  ``randn(shape)`` supplies the standard
    normal draws for delayerr/gainerr (:1655, :1666) instead of the global numpy RNG."""
    ft = NP.float32 if reference_float32 else NP.float64
    ct = NP.complex64 if reference_float32 else NP.complex128
    antpos = NP.asarray(antpos, dtype=NP.float64)
    if antpos.shape[1] == 2:
        antpos = NP.hstack((antpos, NP.zeros((antpos.shape[0], 1))))
    antpos = antpos.astype(ft)                                     # :1593
    nrand = 1
    delays = NP.zeros(antpos.shape[0])
    gains = NP.ones(antpos.shape[0])
    if pointing_info is not None:                                  # :1599-1671
        nrand = pointing_info.get("nrand", 1) or 1
        if pointing_info.get("delays", None) is not None:
            delays = NP.asarray(pointing_info["delays"]).ravel()
        elif "pointing_center" in pointing_info and "delays" not in pointing_info:
            if pointing_info["pointing_coords"] == "altaz":
                pc = altaz2dircos(NP.asarray(pointing_info["pointing_center"]).reshape(1, -1), units="degrees")
            else:
                pc = NP.asarray(pointing_info["pointing_center"]).reshape(1, -1)
            delays = (NP.dot(antpos, pc.T.astype(ft)) / FCNST.c).ravel()       # :1632
        if pointing_info.get("gains", None) is not None:
            gains = NP.asarray(pointing_info["gains"]).ravel()
        if randn is None:
            randn = NP.random.standard_normal
        if pointing_info.get("delayerr", None) is not None:
            delays = delays.reshape(-1, 1) + pointing_info["delayerr"] * randn((antpos.shape[0], nrand))   # :1655
        if pointing_info.get("gainerr", None) is not None:
            gains = gains.reshape(-1, 1) * 10 ** ((pointing_info["gainerr"] / 10.0) * randn((antpos.shape[0], nrand)))   # :1665-1666
    gains = NP.asarray(gains).astype(ft)                           # :1670
    delays = NP.asarray(delays).astype(ft)                         # :1671
    sky = _skypos_to_dircos(skypos, skycoords).astype(ft)          # :1673-1716
    wavelength = NP.asarray(wavelength, dtype=NP.float64).reshape(-1).astype(ft)   # :1728
    geometric_delays = (-NP.dot(antpos, sky.T) / FCNST.c).astype(ft)[:, :, NP.newaxis, NP.newaxis]   # :1730-1731
    if gains.ndim == 1:
        gains = NP.repeat(gains.reshape(-1, 1), nrand, axis=1)
    if delays.ndim == 1:
        delays = NP.repeat(delays.reshape(-1, 1), nrand, axis=1)
    gains = gains.reshape(antpos.shape[0], 1, 1, nrand).astype(ct)  # :1733
    delays = delays.reshape(antpos.shape[0], 1, 1, nrand)           # :1734
    wl = wavelength.reshape(1, 1, -1, 1)                            # :1735
    retvalue = (geometric_delays + delays).astype(ct)               # :1737-1738
    retvalue = NP.exp(1j * 2 * NP.pi * FCNST.c / wl * retvalue).astype(ct)   # :1742
    retvalue = retvalue * (gains / antpos.shape[0])                 # :1743
    retvalue = NP.sum(retvalue.astype(ct), axis=0)                  # :1744
    if power:
        retvalue = NP.abs(retvalue) ** 2
    return retvalue


def primary_beam_generator(skypos, frequency, telescope, freq_scale="GHz", skyunits="degrees", east2ax1=0.0,
                           pointing_info=None, pointing_center=None, short_dipole_approx=False,
                           half_wave_dipole_approx=False, reference_float32=False, randn=None):
    """primary_beams.py:9-441.  Presets hera/hirax/mwa/mwa_dipole/paper and custom shapes
    delta/dipole/dish/gaussian; vla/gmrt/rect/square are out of scope."""
    frequency = NP.asarray(frequency, dtype=NP.float64)
    scale = {"ghz": 1e9, "mhz": 1e6, "khz": 1e3}.get(str(freq_scale).lower(), 1.0)   # :212-217
    frequency = (frequency * scale).reshape(-1)
    if not isinstance(telescope, dict):
        raise TypeError("telescope must be specified as a dictionary")
    wl = FCNST.c / frequency
    afp_kw = dict(reference_float32=reference_float32, randn=randn)

    def _beamformer(ep):                                            # :385-416 / :286-316
        element_locs = telescope["element_locs"]
        pinfo = {key: pointing_info[key] for key in ("delays", "delayerr", "pointing_center", "pointing_coords",
                                                      "gains", "gainerr", "nrand") if key in pointing_info}
        return array_field_pattern(element_locs, skypos, skycoords=skyunits, pointing_info=pinfo,
                                   wavelength=wl, power=False, **afp_kw)

    def _dipole_orientation():                                      # :250-265 / :326-341
        if ("orientation" in telescope) and ("ocoords" in telescope):
            return NP.asarray(telescope["orientation"]).reshape(1, -1), telescope["ocoords"]
        if ("orientation" not in telescope) and ("ocoords" in telescope):
            if telescope["ocoords"] == "altaz":
                return NP.asarray([0.0, 90.0]).reshape(1, -1), "altaz"
            return NP.asarray([1.0, 0.0, 0.0]).reshape(1, -1), "dircos"
        if ("orientation" in telescope) and ("ocoords" not in telescope):
            raise KeyError('key "ocoords" in telescope dictionary not specified.')
        return NP.asarray([1.0, 0.0, 0.0]).reshape(1, -1), "dircos"

    if "id" in telescope:
        tid = telescope["id"]
        if tid in ("hera", "hirax"):                                # :239-247
            dish_dia = 14.0 if tid == "hera" else 6.0
            pb = airy_disk_pattern(dish_dia, skypos, frequency, skyunits=skyunits, peak=1.0,
                                   pointing_center=NP.asarray(telescope["orientation"]),
                                   pointing_coords=telescope["ocoords"], power=True, small_angle_tol=1e-10)
        elif tid == "mwa":                                          # :248-319
            orientation, ocoords = _dipole_orientation()
            ep = dipole_field_pattern(0.74, skypos, dipole_coords=ocoords, dipole_orientation=orientation,
                                      skycoords=skyunits, wavelength=wl, short_dipole_approx=short_dipole_approx,
                                      half_wave_dipole_approx=half_wave_dipole_approx, power=False)[:, :, NP.newaxis]
            if pointing_info is None:                               # :273-285
                pc = NP.asarray([90.0, 270.0]).reshape(1, -1) if skyunits == "altaz" else NP.asarray([0.0, 0.0, 1.0]).reshape(1, -1)
                irap = isotropic_radiators_array_field_pattern(4, 4, 1.1, 1.1, skypos, wl, east2ax1=east2ax1,
                                                               pointing_center=pc, skycoords=skyunits, power=False)[:, :, NP.newaxis]
            else:                                                   # :287-316
                if "element_locs" not in telescope:
                    xlocs, ylocs = NP.meshgrid(1.1 * NP.linspace(-1.5, 1.5, 4), 1.1 * NP.linspace(1.5, -1.5, 4))
                    telescope = dict(telescope)
                    telescope["element_locs"] = NP.hstack((xlocs.reshape(-1, 1), ylocs.reshape(-1, 1), NP.zeros(xlocs.size).reshape(-1, 1)))
                irap = _beamformer(ep)
            pb = NP.mean(NP.abs(ep * irap) ** 2, axis=2)            # :317
        elif tid in ("mwa_dipole", "paper"):                        # :320-349
            dipole_size = 0.74 if tid == "mwa_dipole" else 2.0
            orientation, ocoords = _dipole_orientation()
            ep = dipole_field_pattern(dipole_size, skypos, dipole_coords=ocoords, dipole_orientation=orientation,
                                      skycoords=skyunits, wavelength=wl, short_dipole_approx=short_dipole_approx,
                                      half_wave_dipole_approx=half_wave_dipole_approx, power=False)
            pb = NP.abs(ep) ** 2
        else:
            raise ValueError("preset out of scope for the oracle: {0}".format(tid))
    else:                                                           # :354-416
        shape = telescope.get("shape", "delta")
        nsrc = NP.asarray(skypos).shape[0]
        if shape == "delta":
            ep = 1.0
        elif shape == "dipole":
            ep = dipole_field_pattern(telescope["size"], skypos, dipole_coords=telescope["ocoords"],
                                      dipole_orientation=telescope["orientation"], skycoords=skyunits, wavelength=wl,
                                      short_dipole_approx=short_dipole_approx,
                                      half_wave_dipole_approx=half_wave_dipole_approx, power=False)[:, :, NP.newaxis]
        elif shape == "dish":
            ep = airy_disk_pattern(telescope["size"], skypos, frequency, skyunits=skyunits, peak=1.0,
                                   pointing_center=pointing_center, power=False, small_angle_tol=1e-10)[:, :, NP.newaxis]
        elif shape == "gaussian":
            ep = gaussian_beam(telescope["size"], skypos, frequency, skyunits=skyunits,
                               pointing_center=pointing_center, power=False)[:, :, NP.newaxis]
        else:
            raise ValueError("shape out of scope for the oracle: {0}".format(shape))
        if (pointing_info is not None) and ("element_locs" in telescope):
            irap = _beamformer(ep)
        else:
            irap = NP.ones((nsrc, frequency.size, 1))
        pb = NP.mean(NP.abs(ep * irap) ** 2, axis=2)                # :416

    if "groundplane" in telescope:                                  # :418-439
        gp = 1.0
        if telescope["groundplane"] is not None:
            if ("shape" not in telescope) or (telescope["shape"] != "dish"):
                gp = ground_plane_field_pattern(telescope["groundplane"], skypos, skycoords=skyunits, wavelength=wl,
                                                angle_units="degrees", modifier=telescope.get("ground_modify", None),
                                                power=False)
        pb = pb * gp ** 2
    return pb


# --------------------------------------------------------------------------------------------
# [AU-memory] catalog.SkyModel.generate_spectrum for spec_type='func', name='power-law'
# --------------------------------------------------------------------------------------------
def power_law_spectrum(flux_scale, power_law_index, freq_ref, frequency, flux_offset=None):
    """SM.SkyModel.generate_spectrum [AU-memory] (call site interferometry.py:6249) with the
    parameters built at run_prisim.py:1629-1636: S[s,f] = offset + scale (f/f_ref)^index."""
    frequency = NP.asarray(frequency, dtype=NP.float64).reshape(1, -1)
    flux_scale = NP.asarray(flux_scale, dtype=NP.float64).reshape(-1, 1)
    index = NP.asarray(power_law_index, dtype=NP.float64).reshape(-1, 1)
    freq_ref = NP.broadcast_to(NP.asarray(freq_ref, dtype=NP.float64).reshape(-1, 1), flux_scale.shape)
    out = flux_scale * (frequency / freq_ref) ** index
    if flux_offset is not None:
        out = out + NP.asarray(flux_offset, dtype=NP.float64).reshape(-1, 1)
    return out


# --------------------------------------------------------------------------------------------
# interferometry.py: observe() core
# --------------------------------------------------------------------------------------------
def roi_select(skypos_altaz, roi_radius=None):
    """interferometry.py:6204-6216 (roi_center='zenith'): indices with alt >= 90 - roi_radius."""
    if roi_radius is None:
        roi_radius = 90.0
    m2 = NP.arange(skypos_altaz.shape[0])
    return m2[NP.where(skypos_altaz[:, 0] >= 90.0 - roi_radius)]


def source_taper(baseline_lengths, geometric_delays, channels, src_shape):
    """interferometry.py:6258-6283: vis_wts [nsrc, nbl, nchan].  The reference takes the square
    root of a quantity that can round slightly negative (NaN, :6265); it is clamped at zero here
    (SURVEY.md Appendix C #3)."""
    wl = FCNST.c / NP.asarray(channels, dtype=NP.float64)
    arg = baseline_lengths.reshape(1, -1, 1) ** 2 - (FCNST.c * geometric_delays[:, :, NP.newaxis]) ** 2
    psf = NP.sqrt(NP.maximum(arg, 0.0)) / wl.reshape(1, 1, -1)      # :6265
    src_FWHM = NP.sqrt(src_shape[:, 0] * src_shape[:, 1])           # :6267
    src_FWHM_dircos = 2.0 * NP.sin(0.5 * NP.radians(src_FWHM)).reshape(-1, 1)   # :6268
    sigma = 1.0 / NP.sqrt(2.0 * NP.log(2.0)) / src_FWHM_dircos      # :6270
    return NP.exp(-0.5 * (psf / sigma[:, :, NP.newaxis]) ** 2)      # :6283


def skyvis_snapshot(baselines_enu, skypos_altaz_roi, pbfluxes, channels, pc_altaz, src_shape=None,
                    max_slab_bytes=2.0e8, gradient=False):
    """The phase-sum DFT of one snapshot, interferometry.py:6155-6165 (phase-centre delays),
    :6255 (geometric delays), :6332 + :6340 (phase matrix, sum over sources), evaluated over
    source slabs like the reference's low-memory branch :6348-6376.

    baselines_enu [nbl,3] m, skypos_altaz_roi [nsrc,2] deg, pbfluxes [nsrc,nchan], channels [nchan]
    Hz, pc_altaz [2] deg.  Returns complex128 [nbl, nchan]; with ``gradient=True`` also the
    visibility gradient w.r.t. the baseline vector, G[i,b,f] = sum_s dircos_s[i] * (term of V)
    (gradient_mode='baseline', :6338 / :6343), complex128 [3, nbl, nchan]."""
    baselines_enu = NP.asarray(baselines_enu, dtype=NP.float64)
    channels = NP.asarray(channels, dtype=NP.float64)
    nbl, nchan = baselines_enu.shape[0], channels.size
    skyvis = NP.zeros((nbl, nchan), dtype=NP.complex128)
    nsrc = skypos_altaz_roi.shape[0]
    grad = NP.zeros((3, nbl, nchan), dtype=NP.complex128)
    if nsrc == 0:                                                   # :6378-6382
        return (skyvis, grad) if gradient else skyvis
    skypos_dircos_roi = altaz2dircos(skypos_altaz_roi, "degrees")   # :6263
    pc_dircos = altaz2dircos(pc_altaz, "degrees")                   # :6164
    pc_delay_offsets = geometric_delay(baselines_enu, pc_dircos, altaz=False, hadec=False, dircos=True)   # :6165
    geometric_delays = geometric_delay(baselines_enu, skypos_altaz_roi, altaz=True, hadec=False)          # :6255
    vis_wts = None
    baseline_lengths = NP.sqrt(NP.sum(baselines_enu ** 2, axis=1))
    step = max(1, int(max_slab_bytes / (16.0 * nbl * nchan)))       # :6350-6351 analogue
    for i0 in range(0, nsrc, step):
        sl = slice(i0, min(i0 + step, nsrc))
        phase_matrix = 2.0 * NP.pi * (geometric_delays[sl, :, NP.newaxis] - pc_delay_offsets.reshape(1, -1, 1)) * channels.reshape(1, 1, -1)   # :6332
        term = pbfluxes[sl, NP.newaxis, :] * NP.exp(-1j * phase_matrix)     # :6340
        if src_shape is not None:
            vis_wts = source_taper(baseline_lengths, geometric_delays[sl], channels, src_shape[sl])
            term = term * vis_wts                                   # :6335
        skyvis += NP.sum(term, axis=0)
        if gradient:                                                # :6338, :6343
            grad += NP.sum(skypos_dircos_roi[sl, :, NP.newaxis, NP.newaxis] * term[:, NP.newaxis, :, :], axis=0)
    return (skyvis, grad) if gradient else skyvis


def roi_append_settings(skypos, skycoords, latitude, freq, telescope, roi_info, pinfo=None, lst=None):
    """One call of ROI_parameters.append_settings, interferometry.py:4221-4617, for sky positions in 'hadec',
    'altaz' or 'dircos' (degrees).  Returns (ind, pbeam [n_roi, nchan], radius, center_altaz [1,2] or None).
    * 'ind' given (:4451-4453): used as is; 'pbeam' given with it: cast to float32 (:4466).
    * otherwise radius clamped to [0, 90], default 90 (:4497-4501); centre default zenith, converted to alt-az
      (:4503-4523); zenith-centred (within 1e-2 deg, :4546-4547) selects alt >= 90 - radius (:4551), else
      GEOM.spherematch within radius of the centre (:4549) [AU-memory], called with (alt, az) in the (lon, lat) slots.
    * beam: primary_beam_generator at the ROI positions for all channels if roi_info['pbeam_chromaticity'], else at
      pbeam_reffreq (default centre channel) broadcast over the channels (:4578-4615)."""
    freq = NP.asarray(freq, dtype=NP.float64).ravel()
    skypos = NP.asarray(skypos, dtype=NP.float64)
    altaz = {"hadec": lambda: hadec2altaz(skypos, latitude, units="degrees"), "altaz": lambda: skypos,
             "dircos": lambda: dircos2altaz(skypos, units="degrees")}[skycoords]()
    roi_info = dict(roi_info)
    radius, center = None, None
    if roi_info.get("ind", None) is not None:
        ind = NP.asarray(roi_info["ind"])
        radius = roi_info.get("radius", None)
        if roi_info.get("pbeam", None) is not None and ind.size > 0:
            pb = NP.asarray(roi_info["pbeam"]).reshape(-1, freq.size)
            if ind.size != pb.shape[0]:
                raise ValueError('Number of elements in values in key "ind" and number of rows of values in key "pbeam" must be identical.')
            return ind, pb.astype(NP.float32), radius, center
    else:
        radius = 90.0 if roi_info.get("radius", None) is None else max(0.0, min(roi_info["radius"], 90.0))
        if roi_info.get("center", None) is None:
            center = NP.asarray([90.0, 270.0]).reshape(1, -1)
        else:
            c = NP.asarray(roi_info["center"], dtype=NP.float64).reshape(1, -1)
            cc = roi_info["center_coords"]
            if cc == "dircos":
                center = dircos2altaz(c, units="degrees")
            elif cc == "altaz":
                center = c
            elif cc == "hadec":
                center = hadec2altaz(c, latitude, units="degrees")
            elif cc == "radec":
                if lst is None:
                    raise KeyError("LST not provided for coordinate conversion")
                center = hadec2altaz(NP.asarray([lst - c[0, 0], c[0, 1]]).reshape(1, -1), latitude, units="degrees")
            else:
                raise ValueError("Invalid coordinate system specified for center")
        if sphdist(center[0, 1], center[0, 0], 270.0, 90.0) > 1e-2:
            # the reference hands (alt, az) to spherematch(lon1, lat1, lon2, lat2, ...) in that order (:4549), i.e. the
            # altitude plays the longitude; reproduced as written
            ind = NP.where(sphdist(center[0, 0], center[0, 1], altaz[:, 0], altaz[:, 1]) <= radius)[0]
        else:
            ind = NP.where(altaz[:, 0] >= 90.0 - radius)[0]
    if ind.size == 0:
        return ind, NP.asarray([]), radius, center
    pinfo = dict(pinfo) if pinfo is not None else None
    if pinfo is None:
        raise ValueError("Pointing info dictionary pinfo must be specified.")
    if pinfo.get("pointing_coords", "altaz") == "radec":
        pc = NP.asarray(pinfo["pointing_center"], dtype=NP.float64).reshape(1, -1)
        pinfo["pointing_center"] = hadec2altaz(NP.asarray([lst - pc[0, 0], pc[0, 1]]).reshape(1, -1), latitude, units="degrees")
        pinfo["pointing_coords"] = "altaz"
    reffreq = roi_info.get("pbeam_reffreq", freq[freq.size // 2])
    fcomp = freq if roi_info.get("pbeam_chromaticity", False) else NP.asarray(reffreq, dtype=NP.float64).reshape(-1)
    pbeam = primary_beam_generator(altaz[ind, :], fcomp, telescope, freq_scale="Hz", skyunits="altaz", pointing_info=pinfo)
    return ind, pbeam.astype(NP.float64) * NP.ones(freq.size).reshape(1, -1), radius, center


def uniq_baselines(baseline_locations, redundant=None):
    """interferometry.py:1373-1463: unique baselines by (length to 0.01 m, zenith angle and orientation folded into
    [0, 180) deg to 0.001 arcsec), keyed as formatted strings and ordered as NP.unique orders those strings (:1432-1434).
    redundant=None: all unique baselines; True: those occurring more than once; False: exactly once.  Returns
    (baselines [nu,3], first-occurrence indices, counts, list of index lists of all occurrences)
    (the last via NMO.find_all_occurrences_list1_in_list2 [AU-memory])."""
    bl = NP.asarray(baseline_locations, dtype=NP.float64)
    if bl.shape[1] > 3:
        bl = bl[:, :3]
    elif bl.shape[1] < 3:
        bl = NP.hstack((bl, NP.zeros((bl.shape[0], 3 - bl.shape[1]))))
    blo = NP.angle(bl[:, 0] + 1j * bl[:, 1], deg=True)
    blo[blo >= 180.0] -= 180.0
    blo[blo < 0.0] += 180.0
    bll = NP.sqrt(NP.sum(bl ** 2, axis=1))
    blza = NP.degrees(NP.arccos(bl[:, 2] / bll))
    blstr = ["{0[0]:.2f}_{0[1]:.3f}_{0[2]:.3f}".format(lo) for lo in zip(bll, 3.6e3 * blza, 3.6e3 * blo)]   # :1432
    uniq, ind, invind = NP.unique(blstr, return_index=True, return_inverse=True)
    counts = NP.asarray([blstr.count(u) for u in uniq])
    if redundant is None:
        sel = NP.arange(uniq.size)
    elif redundant:
        sel = NP.where(counts > 1)[0]
    else:
        sel = NP.where(counts == 1)[0]
    retind = ind[sel]
    cnt = counts[sel] if redundant is None or redundant else NP.ones(retind.size)
    occ = [NP.where(invind == invind[i])[0].tolist() for i in retind]
    return bl[retind, :], retind, cnt, occ


def duplicate_counts(labels, blgroups):
    """Index logic of InterferometerArray.duplicate_measurements, interferometry.py:6852-6889.
    labels: sequence of (A2, A1) label tuples of the simulated (unique) baselines; blgroups: dict
    key tuple -> sequence of label tuples redundant with it.  A key found only reversed in `labels` is
    used reversed (:6861-6864); a key missing from its own group is prepended (:6866-6870); a label
    without a group counts once (:6883-6885); a label in two groups raises ValueError (:6880-6882).
    Returns (num_list [nbl] ints, expanded list of label tuples), or (None, labels) when the groups
    hold no more baselines than `labels` (:6852-6857: nothing to do)."""
    labels = [tuple(l) for l in labels]
    groups = {tuple(k): [tuple(l) for l in v] for k, v in blgroups.items()}
    if len(labels) >= sum(len(v) for v in groups.values()):
        return None, labels
    for key in list(groups):
        use = key
        if key not in labels:
            if tuple(reversed(key)) not in labels:
                raise KeyError("Input label {0} not found in attribute labels".format(key))
            use = tuple(reversed(key))
            groups.setdefault(use, groups[key])         # the reference looks the reversed key up in blgroups (KeyError there)
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
    return NP.asarray(num_list, dtype=NP.int64), out


def apply_gradients(gradient_baseline, perturbations, channels):
    """First-order perturbed visibilities, interferometry.py:6811-6819:
    dV = -i 2 pi / lambda * sum_i db[..., i, b] G[i, b, f, t].  gradient_baseline [3,nbl,nchan,nsnap],
    perturbations [..., 3, nbl] (metres) -> [..., nbl, nchan, nsnap]."""
    pert = NP.asarray(perturbations, dtype=NP.float64)
    if pert.ndim == 2:
        pert = pert[NP.newaxis, ...]
    inpshape = pert.shape
    pert = pert.reshape(-1, inpshape[-2], inpshape[-1])
    if pert.shape[1] < 3:                                           # :6801-6806 zero-fill missing axes
        pert = NP.concatenate((pert, NP.zeros((pert.shape[0], 3 - pert.shape[1], pert.shape[2]))), axis=1)
    pert = pert[:, :3, :]
    wl = FCNST.c / NP.asarray(channels, dtype=NP.float64)
    out = -1j * 2.0 * NP.pi / wl.reshape(1, 1, -1, 1) * NP.sum(pert[..., NP.newaxis, NP.newaxis] * gradient_baseline[NP.newaxis, ...], axis=1)
    return out.reshape(tuple(inpshape[:-2]) + gradient_baseline.shape[1:])


def observe_snapshot(baselines_enu, channels, skypos, skycoords, latitude, pointing_center, pointing_coords,
                     telescope, flux_scale, spindex, freq_ref, flux_offset=None, src_shape=None,
                     pb_info=None, roi_radius=None, roi_info=None, lst=None, gradient=False):
    """One pass of InterferometerArray.observe (interferometry.py:5874-6410) for a power-law
    sky given in 'hadec' or 'altaz' coordinates (degrees).  Returns (skyvis [nbl,nchan], m2)."""
    skypos = NP.asarray(skypos, dtype=NP.float64)
    if skycoords == "hadec":                                        # :6176-6177
        skypos_altaz = hadec2altaz(skypos, latitude, units="degrees")
    elif skycoords == "altaz":
        skypos_altaz = skypos
    else:
        raise ValueError("oracle accepts hadec or altaz sky coordinates")
    pc = NP.asarray(pointing_center, dtype=NP.float64).ravel()
    if pointing_coords == "hadec":                                  # :6155-6157
        pc_altaz = hadec2altaz(pc, latitude, units="degrees").ravel()
    elif pointing_coords == "radec":                                # :6158-6160
        pc_altaz = hadec2altaz(NP.asarray([lst - pc[0], pc[1]]), latitude, units="degrees").ravel()
    else:
        pc_altaz = pc
    pb = None
    if roi_info is not None:                                        # :6189-6202
        m2 = NP.asarray(roi_info["ind"])
        pb = NP.asarray(roi_info["pbeam"], dtype=NP.float64).reshape(-1, len(channels))
    else:
        m2 = roi_select(skypos_altaz, roi_radius)                   # :6215-6216
    if m2.size == 0:
        return NP.zeros((NP.asarray(baselines_enu).shape[0], len(channels)), dtype=NP.complex128), m2
    skypos_altaz_roi = skypos_altaz[m2, :]                          # :6219
    fluxes = power_law_spectrum(NP.asarray(flux_scale)[m2], NP.asarray(spindex)[m2],
                                NP.asarray(freq_ref)[m2] if NP.ndim(freq_ref) else freq_ref, channels,
                                None if flux_offset is None else NP.asarray(flux_offset)[m2])   # :6249
    if pb is None:                                                  # :6251-6252
        pb = primary_beam_generator(skypos_altaz_roi, NP.asarray(channels) / 1.0e9, skyunits="altaz",
                                    telescope=telescope, pointing_info=pb_info, pointing_center=pc_altaz,
                                    freq_scale="GHz")
    pbfluxes = pb * fluxes                                          # :6254
    shp = None if src_shape is None else NP.asarray(src_shape, dtype=NP.float64)[m2]
    return skyvis_snapshot(baselines_enu, skypos_altaz_roi, pbfluxes, channels, pc_altaz, src_shape=shp, gradient=gradient), m2


# --------------------------------------------------------------------------------------------
# Tsys / noise  (interferometry.py:6026-6086, :6661-6722)
# --------------------------------------------------------------------------------------------
def system_temperature(Tsysinfo, channels, nbl, bpcorrect=None):
    """interferometry.py:6026-6053: Tnet, or Trx + T0 (f/f0)^spindex, broadcast to [nbl,nchan]."""
    channels = NP.asarray(channels, dtype=NP.float64)
    if Tsysinfo.get("Tnet", None) is not None:
        Tsys = NP.asarray(Tsysinfo["Tnet"], dtype=NP.float64)
        if Tsys.ndim == 0:
            Tsys = Tsys + NP.zeros((nbl, channels.size))
        elif Tsys.size == channels.size:
            Tsys = NP.repeat(Tsys.reshape(1, -1), nbl, axis=0)
        elif Tsys.size == nbl:
            Tsys = NP.repeat(Tsys.reshape(-1, 1), channels.size, axis=1)
        else:
            Tsys = Tsys.reshape(nbl, channels.size)
    else:
        Tsys = Tsysinfo["Trx"] + Tsysinfo["Tant"]["T0"] * (channels / Tsysinfo["Tant"]["f0"]) ** Tsysinfo["Tant"]["spindex"]   # :6034
        Tsys = Tsys.reshape(1, -1) + NP.zeros(nbl).reshape(-1, 1)   # :6037
    if bpcorrect is not None:
        Tsys = Tsys * NP.asarray(bpcorrect).reshape(1, -1)          # :6043-6053
    return Tsys


def thermal_noise_rms(Tsys, A_eff, eff_Q, t_acc, freq_resolution, flux_unit="JY"):
    """interferometry.py:6676-6689.  Tsys [nbl,nchan,nsnap], A_eff/eff_Q [nbl,nchan] or
    scalars, t_acc [nsnap]."""
    Tsys = NP.asarray(Tsys, dtype=NP.float64)
    eff_Q = NP.asarray(eff_Q, dtype=NP.float64)
    A_eff = NP.asarray(A_eff, dtype=NP.float64)
    if eff_Q.ndim == 2:
        eff_Q = eff_Q[:, :, NP.newaxis]
    if A_eff.ndim == 2:
        A_eff = A_eff[:, :, NP.newaxis]
    t_acc = NP.asarray(t_acc, dtype=NP.float64)[NP.newaxis, NP.newaxis, :]
    if flux_unit.upper() == "JY":
        return 2.0 * FCNST.k / NP.sqrt(t_acc * freq_resolution) * (Tsys / A_eff / eff_Q) / Jy   # :6687
    return 1 / NP.sqrt(t_acc * freq_resolution) * Tsys / eff_Q                                   # :6689


def noise_from_normals(vis_rms_freq, nre, nim):
    """interferometry.py:6693 with the two standard-normal draws passed in."""
    return vis_rms_freq / NP.sqrt(2.0) * (nre + 1j * nim)


def add_noise(skyvis_freq, vis_noise_freq, gains=1.0):
    """interferometry.py:6722."""
    return gains * skyvis_freq + vis_noise_freq


# --------------------------------------------------------------------------------------------
# [AU-memory] astroutils.DSP_modules pieces used by delay_transform
# --------------------------------------------------------------------------------------------
def spectral_axis(length, delx=1.0, shift=True):
    """DSP.spectral_axis [AU-memory] (interferometry.py:8114): fftfreq, optionally fftshifted."""
    ax = NP.fft.fftfreq(length, d=delx)
    return NP.fft.fftshift(ax) if shift else ax


def FT1D(inp, ax=-1, inverse=False, shift=False):
    """DSP.FT1D [AU-memory] (interferometry.py:8116-8127): (i)fft along ax then fftshift along ax."""
    out = NP.fft.ifft(inp, axis=ax) if inverse else NP.fft.fft(inp, axis=ax)
    return NP.fft.fftshift(out, axes=ax) if shift else out


def downsampler(inp, factor, axis=-1):
    """DSP.downsampler(method='interp', kind='linear') [AU-memory] (interferometry.py:8131-8134):
    linear interpolation at positions arange(0, n, factor)."""
    inp = NP.asarray(inp)
    n = inp.shape[axis]
    f = interpolate.interp1d(NP.arange(n), inp, kind="linear", axis=axis)
    return f(NP.arange(0, n, factor))


def windowing(N, shape="rect", pad_width=0, centering=True, area_normalize=False, peak=1.0, power_normalize=False):
    """DSP.windowing [AU-memory] (run_prisim.py:954): 4-term Blackman-Harris ('bhw') or
    Blackman-Nuttall ('bnw') window over n/(N-1), zero padding on both sides, normalisation."""
    n = NP.arange(N)
    if shape == "rect":
        win = NP.ones(N)
    else:
        a = {"bhw": (0.35875, 0.48829, 0.14128, 0.01168), "bnw": (0.3635819, 0.4891775, 0.1365995, 0.0106411)}[shape]
        x = 2 * NP.pi * n / (N - 1)
        win = a[0] - a[1] * NP.cos(x) + a[2] * NP.cos(2 * x) - a[3] * NP.cos(3 * x)
    if area_normalize:
        win = win / NP.sum(win)
    elif power_normalize:
        win = win / NP.sqrt(NP.sum(win ** 2))
    else:
        win = win * peak / NP.amax(win)
    if pad_width > 0:
        win = NP.pad(win, (pad_width, pad_width), mode="constant")
    return win


def broadcast_freq_wts(freq_wts, nbl, nchan, n_acc):
    """interferometry.py:8096-8107."""
    freq_wts = NP.asarray(freq_wts)
    if freq_wts.size == nchan:
        return NP.repeat(NP.expand_dims(NP.repeat(freq_wts.reshape(1, -1), nbl, axis=0), axis=2), n_acc, axis=2)
    if freq_wts.size == nchan * n_acc:
        return NP.repeat(NP.expand_dims(freq_wts.reshape(nchan, -1), axis=0), nbl, axis=0)
    if freq_wts.size == nchan * nbl:
        return NP.repeat(NP.expand_dims(freq_wts.reshape(-1, nchan), axis=2), n_acc, axis=2)
    if freq_wts.size == nchan * nbl * n_acc:
        return freq_wts.reshape(nbl, nchan, n_acc)
    raise ValueError("window shape dimensions incompatible with number of channels and/or number of timestamps.")


def delay_transform(vis_freq, bp, bp_wts, freq_resolution, pad=1.0, downsample=True):
    """interferometry.py:8114-8134 (and delay_spectrum.py:1305-1327) for one product.
    vis_freq/bp/bp_wts are [nbl, nchan, nsnap]; returns (lag spectrum, lags)."""
    nchan = vis_freq.shape[1]
    if pad < 0.0:
        pad = 0.0
    if pad == 0.0:                                                  # :8115-8119
        out = FT1D(vis_freq * bp * bp_wts, ax=1, inverse=True, shift=True) * nchan * freq_resolution
        lags = spectral_axis(nchan, delx=freq_resolution, shift=True)
        return out, lags
    npad = int(nchan * pad)                                         # :8123
    out = FT1D(NP.pad(vis_freq * bp * bp_wts, ((0, 0), (0, npad), (0, 0)), mode="constant"), ax=1, inverse=True,
               shift=True) * (npad + nchan) * freq_resolution      # :8124-8127
    lags = spectral_axis(int(nchan * (1 + pad)), delx=freq_resolution, shift=True)   # delay_spectrum.py:1305
    if downsample:
        out = downsampler(out, 1 + pad, axis=1)                     # :8131-8134
        lags = downsampler(lags, 1 + pad)                           # delay_spectrum.py:1327
    return out, lags


def window_N2width(shape="rect"):
    """DSP.window_N2width(n_window=None, shape) [AU-memory]: width of the equivalent rectangular window as a fraction
    of the window length, sum(w / max w) / N on a long window (rect 1, bhw 0.35875, bnw 0.3635819)."""
    w = windowing(1000000, shape=shape.lower())
    return NP.sum(w / w.max()) / w.size


def multi_window_weights(channels, bw_eff, freq_center=None, shape=None):
    """The sub-band windows of multi_window_delay_transform, interferometry.py:8199-8265: [nwin, nchan].
    n_window = round(bw_eff / (frac_width * df)) samples (:8236-8238) of the window shape, centred on the channel
    nearest each frequency centre (LKP.find_1NN within df/2 [AU-memory], sorted by channel :8246-8251), clipped to
    the band and zero elsewhere (:8253-8265)."""
    channels = NP.asarray(channels, dtype=NP.float64)
    df = channels[1] - channels[0]
    bw_eff = NP.asarray(bw_eff, dtype=NP.float64).reshape(-1)
    if NP.any(bw_eff <= 0.0):
        raise ValueError("All values in effective bandwidth must be strictly positive")
    if freq_center is None:
        freq_center = NP.asarray(channels[int(0.5 * channels.size)]).reshape(-1)        # :8210
    else:
        freq_center = NP.asarray(freq_center, dtype=NP.float64).reshape(-1)
        if NP.any((freq_center <= channels.min()) | (freq_center >= channels.max())):
            raise ValueError("Frequency centers must lie strictly inside the observing band")
    if bw_eff.size == 1 and freq_center.size > 1:
        bw_eff = NP.repeat(bw_eff, freq_center.size)
    elif bw_eff.size > 1 and freq_center.size == 1:
        freq_center = NP.repeat(freq_center, bw_eff.size)
    elif bw_eff.size != freq_center.size:
        raise ValueError("Effective bandwidth(s) and frequency center(s) must have same number of elements")
    shape = "rect" if shape is None else shape
    if shape not in ["rect", "bhw", "bnw", "RECT", "BHW", "BNW"]:
        raise ValueError("Invalid value for window shape specified.")
    n_window = NP.round(bw_eff / window_N2width(shape) / df).astype(int)               # :8236-8238
    ind_channels = NP.rint((freq_center - channels[0]) / df).astype(int)                # nearest channel (within df/2)
    order = NP.argsort(ind_channels, kind="stable")                                     # :8247-8251
    ind_channels, n_window = ind_channels[order], n_window[order]
    freq_wts = NP.zeros((ind_channels.size, channels.size))
    for i, ic in enumerate(ind_channels):
        window = windowing(int(n_window[i]), shape=shape.lower(), centering=True)       # :8254
        chan_idx = ic + NP.arange(int(n_window[i])) - int(n_window[i] / 2)              # :8255
        ok = (chan_idx >= 0) & (chan_idx < channels.size)                               # :8256-8262 (out-of-band part dropped)
        freq_wts[i, chan_idx[ok]] = window[ok]
    return freq_wts


def multi_window_delay_transform(vis_freq, bp, freq_wts, freq_resolution, pad=1.0):
    """interferometry.py:8267-8287 for one product: [nbl, nchan, nsnap] -> [nbl, nwin, nchan, nsnap]; also the
    delay-sample correlation length nchan / sum(window) (:8287)."""
    out = []
    for i in range(freq_wts.shape[0]):
        lag, _ = delay_transform(vis_freq, bp, freq_wts[i][NP.newaxis, :, NP.newaxis], freq_resolution, pad=pad)
        out.append(lag)
    return NP.stack(out, axis=1), vis_freq.shape[1] / NP.sum(freq_wts, axis=1)


def subband_weights(channels, bw_eff, freq_center, shape="rect"):
    """Frequency weights of DelaySpectrum.subband_delay_transform, delay_spectrum.py:2153-2176 (fftpow = 1): [nwin, nchan].
    frac_width = DSP.window_N2width(shape, area_normalize=False, power_normalize=True) [AU-memory: sum((w/max w)^2)/N];
    n_window = round(bw_eff / frac_width / df) (:2156-2157); window = sqrt(frac_width n) * power-normalised window (:2166)
    centred on the channel nearest the centre frequency (LKP.find_1NN within df/2), sorted by channel, clipped to the band."""
    channels = NP.asarray(channels, dtype=NP.float64)
    df = channels[1] - channels[0]
    bw_eff = NP.asarray(bw_eff, dtype=NP.float64).reshape(-1)
    freq_center = NP.asarray(freq_center, dtype=NP.float64).reshape(-1)
    w = windowing(1000000, shape=shape.lower())
    frac_width = NP.sum((w / w.max()) ** 2) / w.size
    n_window = NP.round(bw_eff / frac_width / df).astype(int)
    ind = NP.rint((freq_center - channels[0]) / df).astype(int)
    order = NP.argsort(ind, kind="stable")
    ind, n_window = ind[order], n_window[order]
    freq_wts = NP.zeros((ind.size, channels.size))
    for i, ic in enumerate(ind):
        window = NP.sqrt(frac_width * n_window[i]) * windowing(int(n_window[i]), shape=shape.lower(), power_normalize=True)
        k = ic + NP.arange(int(n_window[i])) - int(n_window[i] / 2)
        ok = (k >= 0) & (k < channels.size)
        freq_wts[i, k[ok]] = window[ok]
    return freq_wts


def subband_delay_transform(vis_freq, bp, freq_wts, freq_resolution, pad=1.0):
    """delay_spectrum.py:2178-2201 for one product: [nbl, nchan, nsnap] -> full-resolution [nbl, nwin, nchan + npad, nsnap]
    (no decimation), the lags, and the correlation length nchan / sum(window) (:2201)."""
    nchan = vis_freq.shape[1]
    npad = int(nchan * pad)
    x = vis_freq[:, NP.newaxis, :, :] * bp[:, NP.newaxis, :, :] * freq_wts[NP.newaxis, :, :, NP.newaxis]
    out = FT1D(NP.pad(x, ((0, 0), (0, 0), (0, npad), (0, 0)), mode="constant"), ax=2, inverse=True, shift=True) * (npad + nchan) * freq_resolution
    return out, spectral_axis(nchan + npad, delx=freq_resolution, shift=True), nchan / NP.sum(freq_wts, axis=1)


def subband_resample(lags, lag_kernel, spectra, bw_eff, total_bw):
    """delay_spectrum.py:2220-2240: decimation by min(total_bw / bw_eff); lags and kernel by linear interpolation at
    arange(0, n, factor), spectra by Fourier resampling to round(n / factor) samples (DSP.downsampler(method='FFT') [AU-memory])."""
    from scipy import signal
    factor = NP.min(total_bw / NP.asarray(bw_eff))
    n = lags.size
    pos = NP.arange(0, n, factor)
    f = interpolate.interp1d(NP.arange(n), lag_kernel, kind="linear", axis=2, bounds_error=False, fill_value=NP.nan)
    rlags = NP.interp(pos, NP.arange(n), lags)
    return rlags, f(pos), [signal.resample(x, int(NP.round(n / factor)), axis=2) for x in spectra], (1.0 / NP.asarray(bw_eff)) / (rlags[1] - rlags[0])


# --------------------------------------------------------------------------------------------
# phase centring / projected baselines (interferometry.py:7712-7995)
# --------------------------------------------------------------------------------------------
def enu2xyz(enu, latitude, units="degrees"):
    """GEOM.enu2xyz [AU-memory]: inverse of xyz2enu (interferometry.py:7976)."""
    enu = NP.asarray(enu, dtype=NP.float64).reshape(-1, 3)
    lat = NP.radians(latitude) if units == "degrees" else latitude
    return NP.stack((-NP.sin(lat) * enu[:, 1] + NP.cos(lat) * enu[:, 2], enu[:, 0],
                     NP.cos(lat) * enu[:, 1] + NP.sin(lat) * enu[:, 2]), axis=1)


def phase_rotate(vis, baselines, dircos_current, dircos_new, channels):
    """interferometry.py:7866-7871: vis [nbl,nchan,nsnap] * exp(-2 pi i b.(s_cur - s_new) f / c);
    dircos_* are [nsnap,3]."""
    pos_diff_dircos = NP.asarray(dircos_current) - NP.asarray(dircos_new)
    b_dot_l = NP.dot(NP.asarray(baselines), pos_diff_dircos.T)                               # :7867
    return vis * NP.exp(-1j * 2 * NP.pi * b_dot_l[:, NP.newaxis, :] * NP.asarray(channels).reshape(1, -1, 1) / FCNST.c)


def project_baselines(baselines, ha_deg, dec_deg, latitude):
    """interferometry.py:7973-7985: [nbl, 3, nsnap] uvw-frame baselines for reference HA/Dec per snapshot."""
    ha, dec = NP.radians(NP.asarray(ha_deg, dtype=NP.float64)).ravel(), NP.radians(NP.asarray(dec_deg, dtype=NP.float64)).ravel()
    eq_baselines = enu2xyz(baselines, latitude, units="degrees")
    rot_matrix = NP.asarray([[NP.sin(ha), NP.cos(ha), NP.zeros(ha.size)],
                             [-NP.sin(dec) * NP.cos(ha), NP.sin(dec) * NP.sin(ha), NP.cos(dec)],
                             [NP.cos(dec) * NP.cos(ha), -NP.cos(dec) * NP.sin(ha), NP.sin(dec)]])
    return NP.dot(eq_baselines, rot_matrix)


# --------------------------------------------------------------------------------------------
# External HEALPix beam (scripts/run_prisim.py:1897-1908); healpy / astroutils restated [AU-memory]
# --------------------------------------------------------------------------------------------
def _hp_ring_info(ring, nside):
    """HEALPix RING scheme: (startpix, ringpix, theta, shifted) of ring index 1..4nside-1 (arrays)."""
    ring = NP.asarray(ring, dtype=NP.int64)
    npix, ncap = 12 * nside * nside, 2 * nside * (nside - 1)
    fact2 = 4.0 / npix
    fact1 = 2.0 * nside * fact2
    northring = NP.where(ring > 2 * nside, 4 * nside - ring, ring)
    cap = northring < nside
    tmp = northring.astype(NP.float64) ** 2 * fact2
    theta = NP.where(cap, NP.arctan2(NP.sqrt(NP.clip(tmp * (2.0 - tmp), 0, None)), 1.0 - tmp),
                     NP.arccos(NP.clip((2.0 * nside - northring) * fact1, -1.0, 1.0)))
    ringpix = NP.where(cap, 4 * northring, 4 * nside)
    shifted = NP.where(cap, True, ((northring - nside) & 1) == 0)
    startpix = NP.where(cap, 2 * northring * (northring - 1), ncap + (northring - nside) * ringpix)
    south = northring != ring
    theta = NP.where(south, NP.pi - theta, theta)
    startpix = NP.where(south, npix - startpix - ringpix, startpix)
    return startpix, ringpix, theta, shifted


def healpix_interp_weights(nside, theta, phi):
    """healpy.get_interp_weights(nside, theta, phi) for the RING scheme [AU-memory of healpy/HEALPix C++
    get_interpol]: the two neighbours in each of the two rings bracketing theta.  Returns (pix [4,n], wgt [4,n])."""
    theta = NP.asarray(theta, dtype=NP.float64).ravel()
    phi = NP.asarray(phi, dtype=NP.float64).ravel() % (2 * NP.pi)
    npix = 12 * nside * nside
    z = NP.cos(theta)
    az = NP.abs(z)
    ir1 = NP.where(az <= 2.0 / 3.0, (nside * (2.0 - 1.5 * z)).astype(NP.int64),
                   NP.where(z > 0, (nside * NP.sqrt(3.0 * (1.0 - az))).astype(NP.int64),
                            4 * nside - (nside * NP.sqrt(3.0 * (1.0 - az))).astype(NP.int64) - 1))
    ir2 = ir1 + 1
    pix = NP.zeros((4, theta.size), dtype=NP.int64)
    wgt = NP.zeros((4, theta.size))
    thetas = []
    for k, ir in enumerate((ir1, ir2)):
        valid = (ir > 0) & (ir < 4 * nside)
        sp, nr, th, sh = _hp_ring_info(NP.where(valid, ir, 1), nside)
        dphi = 2 * NP.pi / nr
        shf = NP.where(sh, 0.5, 0.0)
        tmp = phi / dphi - shf
        i1 = NP.where(tmp < 0, tmp.astype(NP.int64) - 1, tmp.astype(NP.int64))
        w1 = (phi - (i1 + shf) * dphi) / dphi
        i2 = i1 + 1
        i1 = NP.where(i1 < 0, i1 + nr, i1)
        i2 = NP.where(i2 >= nr, i2 - nr, i2)
        pix[2 * k], pix[2 * k + 1] = NP.where(valid, sp + i1, 0), NP.where(valid, sp + i2, 0)
        wgt[2 * k], wgt[2 * k + 1] = NP.where(valid, 1.0 - w1, 0.0), NP.where(valid, w1, 0.0)
        thetas.append(NP.where(valid, th, 0.0))
    theta1, theta2 = thetas
    north, south = ir1 == 0, ir2 == 4 * nside
    mid = ~(north | south)
    wt = NP.where(mid, (theta - theta1) / NP.where(mid, theta2 - theta1, 1.0), 0.0)
    wgt[0] = NP.where(mid, wgt[0] * (1 - wt), wgt[0]); wgt[1] = NP.where(mid, wgt[1] * (1 - wt), wgt[1])
    wgt[2] = NP.where(mid, wgt[2] * wt, wgt[2]); wgt[3] = NP.where(mid, wgt[3] * wt, wgt[3])
    wtn = NP.where(north, theta / NP.where(north, theta2, 1.0), 0.0)
    fac = (1.0 - wtn) * 0.25
    wgt[2] = NP.where(north, wgt[2] * wtn + fac, wgt[2]); wgt[3] = NP.where(north, wgt[3] * wtn + fac, wgt[3])
    wgt[0] = NP.where(north, fac, wgt[0]); wgt[1] = NP.where(north, fac, wgt[1])
    pix[0] = NP.where(north, (pix[2] + 2) & 3, pix[0]); pix[1] = NP.where(north, (pix[3] + 2) & 3, pix[1])
    wts = NP.where(south, (theta - theta1) / NP.where(south, NP.pi - theta1, 1.0), 0.0)
    facs = wts * 0.25
    wgt[0] = NP.where(south, wgt[0] * (1 - wts) + facs, wgt[0]); wgt[1] = NP.where(south, wgt[1] * (1 - wts) + facs, wgt[1])
    wgt[2] = NP.where(south, facs, wgt[2]); wgt[3] = NP.where(south, facs, wgt[3])
    pix[2] = NP.where(south, ((pix[0] + 2) & 3) + npix - 4, pix[2]); pix[3] = NP.where(south, ((pix[1] + 2) & 3) + npix - 4, pix[3])
    return pix, wgt


def healpix_interp_along_axis(indata, theta_phi, inloc_axis, outloc_axis, kind="linear"):
    """OPS.healpix_interp_along_axis [AU-memory] (run_prisim.py:1900): healpy.get_interp_val of every column of
    indata [npix, nin] at (theta, phi), then scipy interp1d along the column axis from inloc_axis to outloc_axis."""
    indata = NP.asarray(indata, dtype=NP.float64)
    nside = int(round(NP.sqrt(indata.shape[0] / 12.0)))
    pix, wgt = healpix_interp_weights(nside, theta_phi[:, 0], theta_phi[:, 1])
    spatial = NP.einsum("kn,knf->nf", wgt, indata[pix])
    inloc, outloc = NP.asarray(inloc_axis, dtype=NP.float64), NP.asarray(outloc_axis, dtype=NP.float64)
    if inloc.size == outloc.size and NP.allclose(inloc, outloc):
        return spatial
    return interpolate.interp1d(inloc, spatial, axis=1, kind=kind, bounds_error=False, fill_value="extrapolate")(outloc)


def external_beam_table(external_beam, beam_freqs, src_altaz_deg, chans_hz, chromatic=True, select_freq=None, kind="cubic"):
    """scripts/run_prisim.py:1897-1908: power beam [nsrc, nchan] from a HEALPix (RING) map [npix, nfreq_b]."""
    theta_phi = NP.hstack((NP.pi / 2 - NP.radians(src_altaz_deg[:, 0]).reshape(-1, 1), NP.radians(src_altaz_deg[:, 1]).reshape(-1, 1)))
    if chromatic:
        interp_logbeam = healpix_interp_along_axis(NP.log10(external_beam), theta_phi, beam_freqs, chans_hz, kind=kind)
    else:
        nearest = NP.argmin(NP.abs(NP.asarray(beam_freqs) - select_freq))
        interp_logbeam = healpix_interp_along_axis(NP.log10(NP.repeat(external_beam[:, nearest].reshape(-1, 1), len(chans_hz), axis=1)),
                                                   theta_phi, chans_hz, chans_hz)
    mx = NP.nanmax(interp_logbeam, axis=0)
    mx[mx <= 0.0] = 0.0
    return 10 ** (interp_logbeam - mx.reshape(1, -1))
