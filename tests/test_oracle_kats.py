"""Analytic known-answer tests that pin the oracle itself."""
import numpy as NP
import pytest
import scipy.constants as FCNST

from oracle import prisim_oracle as O

LAT = -30.7224
FREQS = 150e6 + (NP.arange(64) - 32) * 100e3
BL = NP.array([[14.6, 0.0, 0.0], [0.0, 29.2, 0.0], [-43.8, 25.3, 0.1], [300.0, -200.0, 1.0]])
HERA = {"id": "hera", "orientation": [90.0, 270.0], "ocoords": "altaz"}


def test_geometry_roundtrips():
    rng = NP.random.default_rng(0)
    altaz = NP.stack((rng.uniform(1, 89, 50), rng.uniform(0, 360, 50)), 1)
    dc = O.altaz2dircos(altaz)
    assert NP.allclose(NP.sum(dc ** 2, axis=1), 1.0)
    assert NP.allclose(O.dircos2altaz(dc), altaz, atol=1e-10)
    hadec = O.altaz2hadec(altaz, LAT)
    assert NP.allclose(O.hadec2altaz(hadec, LAT), altaz, atol=1e-9)
    # zenith is (HA=0, Dec=lat); the celestial pole sits at alt=|lat| due south for a southern site
    assert NP.allclose(O.hadec2altaz([0.0, LAT], LAT)[0, 0], 90.0)
    pole = O.hadec2altaz([123.0, -90.0], LAT)[0]
    assert NP.allclose(pole, [abs(LAT), 180.0])
    # a source on the equator at HA=+6h (west) is on the horizon due west
    west = O.hadec2altaz([90.0, 0.0], LAT)[0]
    assert NP.allclose(west, [0.0, 270.0], atol=1e-9)
    assert NP.allclose(O.sphdist(10.0, 0.0, 40.0, 0.0), 30.0)
    assert NP.allclose(O.sphdist(0.0, 90.0, 77.0, 0.0), 90.0)


def test_geometric_delay_matches_definition():
    altaz = NP.array([[90.0, 0.0], [30.0, 90.0]])
    tau = O.geometric_delay(BL, altaz, altaz=True, hadec=False)
    assert tau.shape == (2, 4)
    assert NP.allclose(tau[0], BL[:, 2] / FCNST.c)                       # zenith: only the Up component
    assert NP.allclose(tau[1], (NP.cos(NP.radians(30)) * BL[:, 0] + NP.sin(NP.radians(30)) * BL[:, 2]) / FCNST.c)


def test_single_source_at_phase_centre():
    altaz = NP.array([[90.0, 0.0]])
    pbf = NP.full((1, FREQS.size), 3.5)
    V = O.skyvis_snapshot(BL, altaz, pbf, FREQS, NP.array([90.0, 270.0]))
    assert NP.allclose(V, 3.5 + 0j, atol=1e-12)


def test_single_off_axis_source_phase_and_conjugate_symmetry():
    altaz = NP.array([[60.0, 135.0]])
    pc = NP.array([90.0, 270.0])
    pbf = NP.full((1, FREQS.size), 2.0)
    V = O.skyvis_snapshot(BL, altaz, pbf, FREQS, pc)
    s = O.altaz2dircos(altaz)[0] - O.altaz2dircos(pc)[0]
    expect = 2.0 * NP.exp(-2j * NP.pi * FREQS[None, :] * (BL @ s)[:, None] / FCNST.c)
    assert NP.allclose(V, expect, atol=1e-9)
    Vm = O.skyvis_snapshot(-BL, altaz, pbf, FREQS, pc)
    assert NP.allclose(Vm, NP.conj(V), atol=1e-9)                        # V(-b) = V(b)*


def test_linearity_and_horizon_cull():
    rng = NP.random.default_rng(1)
    n = 40
    hadec = NP.stack((rng.uniform(0, 360, n), NP.degrees(NP.arcsin(rng.uniform(-1, 1, n)))), 1)
    kw = dict(baselines_enu=BL, channels=FREQS, skypos=hadec, skycoords="hadec", latitude=LAT,
              pointing_center=[0.0, LAT], pointing_coords="hadec", telescope=HERA, spindex=NP.full(n, -0.8),
              freq_ref=150e6)
    S = rng.uniform(1, 10, n)
    V1, m2 = O.observe_snapshot(flux_scale=S, **kw)
    V2, _ = O.observe_snapshot(flux_scale=3 * S, **kw)
    assert NP.allclose(V2, 3 * V1)
    altaz = O.hadec2altaz(hadec, LAT)
    assert NP.array_equal(m2, NP.where(altaz[:, 0] >= 0)[0])
    # brighten only below-horizon sources: nothing changes
    S3 = S.copy(); S3[altaz[:, 0] < 0] *= 100
    V3, _ = O.observe_snapshot(flux_scale=S3, **kw)
    assert NP.allclose(V3, V1)
    # empty ROI -> zeros
    V0, m0 = O.observe_snapshot(flux_scale=S, roi_radius=0.0, **kw)
    assert m0.size == 0 and NP.all(V0 == 0)


def test_airy_beam_kats():
    f = NP.array([150e6])
    lam = FCNST.c / f[0]
    assert NP.allclose(O.airy_disk_pattern(14.0, NP.array([[90.0, 0.0]]), f, pointing_center=NP.array([90.0, 270.0]),
                                           pointing_coords="altaz"), 1.0)
    x_null = NP.arcsin(3.8317059702075125 * lam / (NP.pi * 14.0))         # (pi D/lambda) sin x = 3.8317
    pat = O.airy_disk_pattern(14.0, NP.array([[90.0 - NP.degrees(x_null), 33.0]]), f, power=True)
    assert pat[0, 0] < 1e-20
    # below the horizon -> exactly zero; hera preset == airy power with the preset's pointing
    assert O.primary_beam_generator(NP.array([[-1.0, 10.0]]), f / 1e9, HERA, skyunits="altaz")[0, 0] == 0.0
    alt = NP.array([[80.0, 12.0], [45.0, 200.0]])
    assert NP.allclose(O.primary_beam_generator(alt, f / 1e9, HERA, skyunits="altaz"),
                       O.airy_disk_pattern(14.0, alt, f, power=True))
    # 'dish' shape squares the field pattern in the wrapper
    dish = {"shape": "dish", "size": 14.0, "ocoords": "altaz", "orientation": [90.0, 270.0]}
    assert NP.allclose(O.primary_beam_generator(alt, f / 1e9, dish, skyunits="altaz", pointing_center=NP.array([90.0, 270.0])),
                       O.airy_disk_pattern(14.0, alt, f, power=True))


def test_gaussian_and_dipole_and_groundplane_kats():
    f = NP.array([150e6])
    lam = FCNST.c / f[0]
    D = 14.0
    # primary_beams.py:717-724: power = exp(-(sin x / sigma_l)^2), sigma_l = lambda sqrt(2 ln2)/(pi D)
    sig = lam * NP.sqrt(2 * NP.log(2)) / (NP.pi * D)
    x = 3.0
    g = O.gaussian_beam(D, NP.array([[90.0 - x, 0.0]]), f, power=True)[0, 0]
    assert NP.allclose(g, NP.exp(-(NP.sin(NP.radians(x)) / sig) ** 2))
    # short dipole along east: sin(theta) -> zero towards east at the horizon, one at zenith
    sd = O.dipole_field_pattern(0.74, NP.array([[90.0, 0.0], [0.0, 90.0]]), dipole_coords="dircos",
                                dipole_orientation=NP.array([1.0, 0, 0]), skycoords="altaz", wavelength=NP.array([lam]),
                                short_dipole_approx=True, half_wave_dipole_approx=False)
    assert NP.allclose(sd[:, 0], [1.0, 0.0], atol=1e-12)
    # general dipole is normalised to one broadside
    gd = O.dipole_field_pattern(0.74, NP.array([[90.0, 0.0]]), dipole_coords="dircos", dipole_orientation=NP.array([1.0, 0, 0]),
                                skycoords="altaz", wavelength=NP.array([lam]), half_wave_dipole_approx=False)
    assert NP.allclose(gd, 1.0)
    gp = O.ground_plane_field_pattern(0.3, NP.array([[90.0, 0.0], [30.0, 10.0]]), skycoords="altaz", wavelength=NP.array([lam]), power=False)
    k = 2 * NP.pi / lam
    assert NP.allclose(gp[:, 0], [1.0, NP.sin(k * 0.3 * 0.5) / NP.sin(k * 0.3)])


def test_array_factor_analytic_equals_element_sum():
    """The 4x4 closed form (:1467-1473) equals the explicit element sum (:1742-1744) for a
    zenith-phased regular grid."""
    lam = NP.array([FCNST.c / 185e6])
    altaz = NP.array([[70.0, 20.0], [50.0, 250.0], [89.0, 100.0]])
    xl, yl = NP.meshgrid(1.1 * NP.linspace(-1.5, 1.5, 4), 1.1 * NP.linspace(1.5, -1.5, 4))
    locs = NP.hstack((xl.reshape(-1, 1), yl.reshape(-1, 1), NP.zeros((16, 1))))
    ana = O.isotropic_radiators_array_field_pattern(4, 4, 1.1, 1.1, altaz, lam, east2ax1=0.0, pointing_center=NP.array([90.0, 270.0]),
                                                    skycoords="altaz", power=True)
    ele = O.array_field_pattern(locs, altaz, skycoords="altaz", pointing_info=None, wavelength=lam, power=True)[:, :, 0]
    assert NP.allclose(ana, ele, rtol=1e-9, atol=1e-12)
    # the reference's float32 evaluation deviates at the 1e-5..1e-4 level
    ele32 = O.array_field_pattern(locs, altaz, skycoords="altaz", pointing_info=None, wavelength=lam, power=True, reference_float32=True)[:, :, 0]
    assert NP.abs(ele32 - ele).max() < 1e-3


def test_noise_rms_formula_and_statistics():
    Tsys = NP.full((3, 8, 2), 300.0)
    rms = O.thermal_noise_rms(Tsys, 100.0, 0.96, [10.0, 20.0], 100e3)
    expect = 2 * FCNST.k * 300.0 / (100.0 * 0.96 * NP.sqrt(10.0 * 100e3)) / 1e-26
    assert NP.allclose(rms[:, :, 0], expect) and NP.allclose(rms[:, :, 1], expect / NP.sqrt(2))
    rng = NP.random.default_rng(3)
    nz = O.noise_from_normals(NP.full(200000, 2.0), rng.standard_normal(200000), rng.standard_normal(200000))
    assert abs(NP.sqrt(NP.mean(NP.abs(nz) ** 2)) / 2.0 - 1) < 0.01


def test_delay_transform_properties():
    rng = NP.random.default_rng(4)
    nbl, nchan, nt = 3, 64, 2
    df = 100e3
    x = rng.standard_normal((nbl, nchan, nt)) + 1j * rng.standard_normal((nbl, nchan, nt))
    ones = NP.ones((nbl, nchan, nt))
    X0, lags0 = O.delay_transform(x, ones, ones, df, pad=0.0)
    X1, lags1 = O.delay_transform(x, ones, ones, df, pad=1.0)
    assert NP.allclose(X0, X1) and NP.allclose(lags0, lags1)              # pad=1 == pad=0 identity
    # Parseval: sum |X|^2 / (N df)^2 * N = sum |x|^2
    assert NP.allclose(NP.sum(NP.abs(X0) ** 2, axis=1) / (nchan * df ** 2), NP.sum(NP.abs(x) ** 2, axis=1))
    # flat-spectrum source: delay spectrum peaks at tau = b.(s - s_pc)/c inside the horizon limits
    f = 150e6 + (NP.arange(256) - 128) * df
    altaz = NP.array([[50.0, 60.0]]); pc = NP.array([90.0, 270.0])
    bl = NP.array([[250.0, 80.0, 0.0]])
    V = O.skyvis_snapshot(bl, altaz, NP.ones((1, f.size)), f, pc)[:, :, None]
    w = O.windowing(f.size, "bhw", area_normalize=True) * f.size
    L, lags = O.delay_transform(V, NP.ones_like(V.real), w[None, :, None], df, pad=1.0)
    tau = (bl @ (O.altaz2dircos(altaz)[0] - O.altaz2dircos(pc)[0]))[0] / FCNST.c
    # vis ~ exp(-2 pi i f tau) and the ifft kernel is exp(+2 pi i f lag): the peak sits at lag = +tau
    assert abs(lags[NP.argmax(NP.abs(L[0, :, 0]))] - tau) <= 1.0 / (f.size * df)
    lim = O.horizon_delay_limits(bl, O.altaz2dircos(pc)[0])[0]
    assert lim[0] <= tau <= lim[1]
    # generic padding agrees with explicit numpy
    Xh, _ = O.delay_transform(x, ones, ones, df, pad=0.5)
    npad = 32
    ref = NP.fft.fftshift(NP.fft.ifft(NP.pad(x, ((0, 0), (0, npad), (0, 0))), axis=1), axes=1) * (nchan + npad) * df
    pos = NP.arange(0, nchan + npad, 1.5)
    lo = NP.floor(pos).astype(int); fr = pos - lo
    hi = NP.minimum(lo + 1, nchan + npad - 1)
    exp = ref[:, lo, :] * (1 - fr)[None, :, None] + ref[:, hi, :] * fr[None, :, None]
    assert NP.allclose(Xh, exp)


def test_window_and_freq_wts_broadcast():
    w = O.windowing(128, "bhw", area_normalize=True)
    assert NP.allclose(w.sum(), 1.0) and NP.allclose(w, w[::-1]) and w.argmax() in (63, 64)
    assert NP.allclose(O.windowing(16, "rect"), 1.0)
    assert O.broadcast_freq_wts(NP.ones(8), 3, 8, 2).shape == (3, 8, 2)
    assert O.broadcast_freq_wts(NP.ones(24), 3, 8, 2).shape == (3, 8, 2)
    with pytest.raises(ValueError):
        O.broadcast_freq_wts(NP.ones(7), 3, 8, 2)


def test_taper_limits():
    bl = NP.array([[100.0, 0, 0], [1000.0, 0, 0]])
    altaz = NP.array([[90.0, 0.0]])
    f = NP.array([150e6])
    tau = O.geometric_delay(bl, altaz, altaz=True, hadec=False)
    blen = NP.sqrt(NP.sum(bl ** 2, axis=1))
    w_small = O.source_taper(blen, tau, f, NP.array([[1e-6, 1e-6, 0.0]]))
    assert NP.allclose(w_small, 1.0, atol=1e-6)                          # point-like source: no taper
    fwhm = 0.5
    w = O.source_taper(blen, tau, f, NP.array([[fwhm, fwhm, 0.0]]))
    u = blen / (FCNST.c / f[0])
    d = 2 * NP.sin(NP.radians(fwhm) / 2)
    assert NP.allclose(w[0, :, 0], NP.exp(-NP.log(2) * (u * d) ** 2))     # half power at u d = 1


def test_healpix_interpolation_kats():
    """The RING-scheme bilinear interpolation restated for the external-beam path (run_prisim.py:1897-1908)."""
    from prisim_b200.synthetic import healpix_ring_centers
    for nside in (1, 2, 8, 32):
        npix = 12 * nside * nside
        ra, dec = healpix_ring_centers(nside)
        theta_c, phi_c = NP.radians(90.0 - dec), NP.radians(ra)
        rng = NP.random.default_rng(nside)
        theta, phi = NP.arccos(rng.uniform(-1, 1, 4000)), rng.uniform(0, 2 * NP.pi, 4000)
        pix, wgt = O.healpix_interp_weights(nside, theta, phi)
        assert pix.min() >= 0 and pix.max() < npix
        assert NP.allclose(wgt.sum(axis=0), 1.0) and wgt.min() >= -1e-12
        # at a pixel centre the interpolation returns that pixel
        pc, wc = O.healpix_interp_weights(nside, theta_c, phi_c)
        val = NP.arange(npix, dtype=float)
        assert NP.allclose(NP.sum(wc * val[pc], axis=0), val, atol=1e-9 * npix)
        # a smooth function is reproduced to O(pixel size^2)
        fmap = NP.cos(theta_c) + 0.3 * NP.sin(theta_c) * NP.cos(phi_c)
        got = NP.sum(wgt * fmap[pix], axis=0)
        assert NP.abs(got - (NP.cos(theta) + 0.3 * NP.sin(theta) * NP.cos(phi))).max() < 1.5 / nside ** 2 + (0.5 if nside == 1 else 0)
    # the two interpolations commute (what the device path relies on)
    nside = 8
    ra, dec = healpix_ring_centers(nside)
    bf = NP.linspace(100e6, 200e6, 6)
    beam = NP.exp(-((90.0 - dec[:, None]) / (40.0 * 150e6 / bf[None, :])) ** 2) + 1e-4
    altaz = NP.stack((rng.uniform(0, 90, 50), rng.uniform(0, 360, 50)), 1)
    chans = NP.linspace(110e6, 190e6, 9)
    t1 = O.external_beam_table(beam, bf, altaz, chans, kind="cubic")
    from scipy import interpolate
    logmap = interpolate.interp1d(bf, NP.log10(beam), axis=1, kind="cubic")(chans)
    tp = NP.stack((NP.radians(90.0 - altaz[:, 0]), NP.radians(altaz[:, 1])), 1)
    lb = O.healpix_interp_along_axis(logmap, tp, chans, chans)
    mx = NP.clip(lb.max(axis=0), 0, None)
    assert NP.allclose(t1, 10 ** (lb - mx), rtol=1e-12)
    assert t1.max() <= 1.0 + 1e-12
