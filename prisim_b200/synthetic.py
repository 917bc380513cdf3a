"""Synthetic inputs for the five BASELINE.json configurations.

Pure numpy input producers shared by tests, ``bench.py`` and the examples: arrays, catalogues,
channels and observing parameters.  Nothing here evaluates visibilities.  Seeds are
20261017 + configuration index; site is HERA's (lat -30.7224, lon 21.4278,
examples/simparms/defaultparms.yaml:57,67 of the reference).
"""
from __future__ import annotations

import numpy as NP

from .interferometry import baseline_generator, hexagon_generator, orient_and_sort_baselines
from .skymodel import SkyModel

LATITUDE = -30.7224
LONGITUDE = 21.4278
SEED0 = 20261017


def channels(nchan, df, f_center):
    """run_prisim.py:900: chans = f0 + (arange(n) - n/2) * df."""
    return f_center + (NP.arange(nchan) - nchan // 2) * df


def hera_layout(n_side=3, spacing=14.6, n_outriggers=0):
    """Hexagonal core (interferometry.py:958-989) plus deterministic outriggers on two rings
    (300 m / 450 m alternating, azimuth k*360/n): 'HERA-350' = n_side 11 + 19 outriggers."""
    xy, _ = hexagon_generator(spacing, n_side=n_side)
    if n_outriggers > 0:
        k = NP.arange(n_outriggers)
        radius = NP.where(k % 2 == 0, 300.0, 450.0)
        az = NP.radians(k * 360.0 / n_outriggers)
        xy = NP.vstack((xy, NP.stack((radius * NP.sin(az), radius * NP.cos(az)), axis=1)))
    ant = NP.hstack((xy, NP.zeros((xy.shape[0], 1))))
    return ant


def array_baselines(ant, sort=True):
    bl, labels, ids = baseline_generator(ant, ant_label=NP.arange(ant.shape[0]).astype(str), auto=False, conjugate=False)
    if sort:
        bl, labels, _ = orient_and_sort_baselines(bl, labels)
    return bl, labels


def point_source_catalog(nsrc, rng, dec_max=90.0, flux_law="loguniform", smin=0.1, smax=100.0, spindex_mean=-0.83,
                         spindex_rms=0.2, f_ref=150e6, src_fwhm_deg=None):
    """Uniform on the sphere below dec_max; fluxes log-uniform or Euclidean counts dN/dS ~ S^-2.5."""
    ra = rng.uniform(0.0, 360.0, nsrc)
    sin_max = NP.sin(NP.radians(dec_max))
    dec = NP.degrees(NP.arcsin(rng.uniform(-1.0, sin_max, nsrc)))
    if flux_law == "loguniform":
        flux = 10 ** rng.uniform(NP.log10(smin), NP.log10(smax), nsrc)
    else:   # dN/dS ~ S^-2.5  ->  inverse CDF of S^-1.5 between smin and smax
        u = rng.uniform(0.0, 1.0, nsrc)
        flux = (smin ** -1.5 - u * (smin ** -1.5 - smax ** -1.5)) ** (-1.0 / 1.5)
    spindex = rng.normal(spindex_mean, spindex_rms, nsrc)
    parms = {"name": NP.arange(nsrc).astype(str), "frequency": NP.asarray([f_ref]), "location": NP.stack((ra, dec), axis=1),
             "coords": "radec", "epoch": "J2000", "spec_type": "func",
             "spec_parms": {"name": NP.repeat("power-law", nsrc), "power-law-index": spindex,
                            "freq-ref": NP.full(nsrc, f_ref), "flux-scale": flux, "flux-offset": NP.zeros(nsrc)}}
    if src_fwhm_deg is not None:
        fw = NP.broadcast_to(NP.asarray(src_fwhm_deg, dtype=NP.float64), (nsrc,))
        parms["src_shape"] = NP.stack((fw, fw, NP.zeros(nsrc)), axis=1)
    return SkyModel(init_parms=parms)


def healpix_ring_centers(nside):
    """(RA, Dec) in degrees of HEALPix RING-ordered pixel centres (standard closed-form scheme;
    healpy is not available in this image)."""
    npix = 12 * nside * nside
    ncap = 2 * nside * (nside - 1)
    p = NP.arange(npix)
    z = NP.empty(npix)
    phi = NP.empty(npix)
    # north polar cap
    m = p < ncap
    ph = (p[m] + 1) / 2.0
    i = NP.floor(NP.sqrt(ph - NP.sqrt(NP.floor(ph)))).astype(NP.int64) + 1
    j = p[m] + 1 - 2 * i * (i - 1)
    z[m] = 1.0 - i * i / (3.0 * nside * nside)
    phi[m] = NP.pi / (2.0 * i) * (j - 0.5)
    # equatorial belt
    m = (p >= ncap) & (p < npix - ncap)
    pp = p[m] - ncap
    i = pp // (4 * nside) + nside
    j = pp % (4 * nside) + 1
    s = 1.0 + ((i + nside) % 2)            # fodd = 0.5 * s
    z[m] = (2.0 * nside - i) * 2.0 / (3.0 * nside)
    phi[m] = NP.pi / (2.0 * nside) * (j - 0.5 * s)
    # south polar cap
    m = p >= npix - ncap
    ps = npix - p[m]
    ph = ps / 2.0
    i = NP.floor(NP.sqrt(ph - NP.sqrt(NP.floor(ph)))).astype(NP.int64) + 1
    j = 4 * i + 1 - (ps - 2 * i * (i - 1))
    z[m] = -1.0 + i * i / (3.0 * nside * nside)
    phi[m] = NP.pi / (2.0 * i) * (j - 0.5)
    return NP.degrees(phi) % 360.0, NP.degrees(NP.arcsin(NP.clip(z, -1.0, 1.0)))


def diffuse_healpix_catalog(nside, rng, f_ref=150e6):
    """Config 3 sky: T150 = lognormal(ln 200 K, 0.5) * (1 + 5 exp(-(b/10deg)^2)) with b ~ Dec as a
    stand-in for Galactic latitude; S = 2 k T (nu/c)^2 Omega_pix / Jy (run_prisim.py:1220), spectral
    index = beta + 2 with beta ~ N(-2.5, 0.1) (:1223), source FWHM = pixel size (:1230-1231)."""
    import scipy.constants as FCNST
    ra, dec = healpix_ring_centers(nside)
    npix = ra.size
    t150 = rng.lognormal(NP.log(200.0), 0.5, npix) * (1.0 + 5.0 * NP.exp(-(dec / 10.0) ** 2))
    omega = 4.0 * NP.pi / npix
    flux = 2.0 * FCNST.k * t150 * (f_ref / FCNST.c) ** 2 * omega / 1.0e-26
    beta = rng.normal(-2.5, 0.1, npix)
    fwhm = NP.degrees(NP.sqrt(omega))
    parms = {"name": NP.arange(npix).astype(str), "frequency": NP.asarray([f_ref]), "location": NP.stack((ra, dec), axis=1),
             "coords": "radec", "epoch": "J2000", "spec_type": "func",
             "spec_parms": {"name": NP.repeat("power-law", npix), "power-law-index": beta + 2.0,
                            "freq-ref": NP.full(npix, f_ref), "flux-scale": flux, "flux-offset": NP.zeros(npix)},
             "src_shape": NP.stack((NP.full(npix, fwhm), NP.full(npix, fwhm), NP.zeros(npix)), axis=1)}
    return SkyModel(init_parms=parms)


HERA_TELESCOPE = {"id": "hera", "shape": "dish", "size": 14.0, "orientation": [90.0, 270.0], "ocoords": "altaz",
                  "groundplane": None}


def config1(nsrc=1000, nchan=128, nsnap=10, seed=SEED0 + 1):
    """C1: HERA-19, 1k point sources, 128 x 100 kHz at 150 MHz, 10 snapshots of 1080 s, Airy dish."""
    rng = NP.random.default_rng(seed)
    ant = hera_layout(n_side=3)
    bl, labels = array_baselines(ant)
    return {"name": "C1", "ant": ant, "baselines": bl, "labels": labels, "channels": channels(nchan, 100e3, 150e6),
            "skymodel": point_source_catalog(nsrc, rng), "telescope": dict(HERA_TELESCOPE), "latitude": LATITUDE,
            "nsnap": nsnap, "t_acc": 1080.0, "lst0_deg": 0.0, "pointing_hadec": NP.asarray([0.0, LATITUDE])}


def config2(nsrc=300000, nchan=1024, n_side=11, n_outriggers=19, seed=SEED0 + 2):
    """C2: HERA-350 (61,075 baselines) x 1024 x 97.65625 kHz x 300k GLEAM-shaped sources, 1 snapshot."""
    rng = NP.random.default_rng(seed)
    ant = hera_layout(n_side=n_side, n_outriggers=n_outriggers)
    bl, labels = array_baselines(ant)
    sky = point_source_catalog(nsrc, rng, dec_max=30.0, flux_law="euclidean", smin=0.05, smax=50.0)
    return {"name": "C2", "ant": ant, "baselines": bl, "labels": labels, "channels": channels(nchan, 97656.25, 150e6),
            "skymodel": sky, "telescope": dict(HERA_TELESCOPE), "latitude": LATITUDE, "nsnap": 1, "t_acc": 10.7,
            "lst0_deg": 0.0, "pointing_hadec": NP.asarray([0.0, LATITUDE])}


def config3(nside=256, nchan=256, n_side=11, nsnap=100, seed=SEED0 + 3):
    """C3: HEALPix diffuse sky x HERA-331 x 256 x 390.625 kHz, drift scan (taper on)."""
    rng = NP.random.default_rng(seed)
    ant = hera_layout(n_side=n_side)
    bl, labels = array_baselines(ant)
    return {"name": "C3", "ant": ant, "baselines": bl, "labels": labels, "channels": channels(nchan, 390625.0, 150e6),
            "skymodel": diffuse_healpix_catalog(nside, rng), "telescope": dict(HERA_TELESCOPE), "latitude": LATITUDE,
            "nsnap": nsnap, "t_acc": 108.0, "lst0_deg": 0.0, "pointing_hadec": NP.asarray([0.0, LATITUDE])}


def config4(ntiles=128, nsrc=50000, nchan=768, seed=SEED0 + 4):
    """C4: MWA Phase-II-like 128 tiles (N(0, 300 m) truncated at 1.5 km), 4x4 dipole tile beam with
    435 ps-quantised pointing delays (run_prisim.py:585), ground plane 0.3 m, 768 x 40 kHz at 185 MHz."""
    import scipy.constants as FCNST
    rng = NP.random.default_rng(seed)
    pos = NP.empty((0, 2))
    while pos.shape[0] < ntiles:
        cand = rng.normal(0.0, 300.0, (2 * ntiles, 2))
        pos = NP.vstack((pos, cand[NP.hypot(cand[:, 0], cand[:, 1]) < 1500.0]))
    ant = NP.hstack((pos[:ntiles], NP.zeros((ntiles, 1))))
    bl, labels = array_baselines(ant)
    xlocs, ylocs = NP.meshgrid(1.1 * NP.linspace(-1.5, 1.5, 4), 1.1 * NP.linspace(1.5, -1.5, 4))
    element_locs = NP.hstack((xlocs.reshape(-1, 1), ylocs.reshape(-1, 1), NP.zeros((16, 1))))
    pc_altaz = NP.asarray([52.806, 101.31])
    alt, az = NP.radians(pc_altaz)
    pc_dircos = NP.asarray([NP.cos(alt) * NP.sin(az), NP.cos(alt) * NP.cos(az), NP.sin(alt)])
    delays = NP.dot(element_locs, pc_dircos) / FCNST.c
    delays = NP.round((delays - delays.min()) / 435e-12) * 435e-12
    telescope = {"id": "mwa", "shape": "dipole", "size": 0.74, "orientation": [1.0, 0.0, 0.0], "ocoords": "dircos",
                 "groundplane": 0.3, "element_locs": element_locs}
    sky = point_source_catalog(nsrc, rng, dec_max=30.0, flux_law="euclidean", smin=0.05, smax=50.0, f_ref=185e6)
    return {"name": "C4", "ant": ant, "baselines": bl, "labels": labels, "channels": channels(nchan, 40e3, 185e6),
            "skymodel": sky, "telescope": telescope, "latitude": -26.701, "nsnap": 1, "t_acc": 8.0, "lst0_deg": 0.0,
            "pb_info": {"delays": delays, "pointing_center": pc_altaz, "pointing_coords": "altaz"},
            "pointing_altaz": pc_altaz}


def config5(nsnap=1000, **kw):
    """C5: C2 array x nsnap snapshots of 10.7 s + Tsys noise + windowed delay transform."""
    cfg = config2(**kw)
    cfg.update(name="C5", nsnap=nsnap, t_acc=10.7,
               Tsysinfo={"Trx": 50.0, "Tant": {"T0": 200.0, "f0": 150e6, "spindex": -2.55}, "Tnet": None},
               A_eff=154.0 * 0.65, eff_Q=0.96)
    return cfg
