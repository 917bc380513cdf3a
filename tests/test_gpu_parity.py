"""GPU parity tests (run with `-m gpu` on the B200 box): every kernel behind the C-ABI against
the oracle and against the golden vectors generated from the reference's own code.

Tolerance (BASELINE.json north_star): max |dV| <= 1e-5 * rms(V) per baseline (rms over the
baseline's channels and snapshots); same for delay spectra.  Index work
(ROI selection) is exact."""
import os

import numpy as NP
import pytest
import torch

from oracle import prisim_oracle as O

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5
LAT = -30.7224


def rel_err_per_baseline(Vg, Vo):
    axes = tuple(range(1, Vo.ndim))
    rms_b = NP.sqrt(NP.mean(NP.abs(Vo) ** 2, axis=axes, keepdims=True))
    rms_b = NP.where(rms_b > 0, rms_b, 1.0)
    return float((NP.abs(Vg - Vo) / rms_b).max())


@pytest.fixture(scope="module")
def eng():
    from prisim_b200 import engine
    return engine


def _random_sky(rng, n):
    ha = rng.uniform(0, 360, n)
    dec = NP.degrees(NP.arcsin(rng.uniform(-1, 1, n)))
    return NP.stack((ha, dec), 1)


def _hex_bl(scale=1.0):
    from prisim_b200 import synthetic as S
    return S.array_baselines(S.hera_layout(3))[0] * scale


# ------------------------------------------------------------------ cull
@pytest.mark.parametrize("n", [0, 1, 31, 257, 5000])
def test_sky_cull_exact(eng, n):
    rng = NP.random.default_rng(n)
    hadec = _random_sky(rng, n)
    dircos, idx = eng.sky_cull(hadec, "hadec", latitude_deg=LAT)
    altaz = O.hadec2altaz(hadec, LAT) if n else NP.zeros((0, 2))
    m2 = O.roi_select(altaz) if n else NP.zeros(0, dtype=int)
    assert NP.array_equal(idx.cpu().numpy(), m2)
    if m2.size:
        assert NP.abs(dircos.cpu().numpy() - O.altaz2dircos(altaz[m2])).max() < 1e-14


def test_sky_cull_roi_radius_and_centre_and_coords(eng):
    rng = NP.random.default_rng(5)
    hadec = _random_sky(rng, 3000)
    altaz = O.hadec2altaz(hadec, LAT)
    for radius in (20.0, 60.0, 90.0, 180.0):
        _, idx = eng.sky_cull(hadec, "hadec", latitude_deg=LAT, roi_radius_deg=radius)
        assert NP.array_equal(idx.cpu().numpy(), O.roi_select(altaz, radius))
    d1, i1 = eng.sky_cull(altaz, "altaz")
    d2, i2 = eng.sky_cull(O.altaz2dircos(altaz), "dircos")
    assert torch.equal(i1, i2) and (d1 - d2).abs().max().item() < 1e-15
    pc = O.altaz2dircos([60.0, 100.0])[0]
    _, ic = eng.sky_cull(altaz, "altaz", roi_radius_deg=25.0, roi_center_dircos=pc)
    sep = O.sphdist(altaz[:, 1], altaz[:, 0], 100.0, 60.0)
    assert NP.array_equal(ic.cpu().numpy(), NP.where(sep <= 25.0)[0])


# ------------------------------------------------------------------ beams / amplitude table
BEAM_CASES = {
    "hera": ({"id": "hera", "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz"}, {}),
    "hirax_offzenith": ({"id": "hirax", "orientation": NP.asarray([75.0, 120.0]), "ocoords": "altaz"}, {}),
    "mwa_dipole_gp": ({"id": "mwa_dipole", "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3}, {}),
    "paper": ({"id": "paper", "orientation": NP.asarray([0.0, 90.0]), "ocoords": "altaz"}, {}),
    "mwa_analytic": ({"id": "mwa", "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3}, {}),
    "delta": ({"shape": "delta"}, {}),
    "dish": ({"shape": "dish", "size": 14.0, "ocoords": "altaz", "orientation": NP.asarray([90.0, 270.0])}, {"pc": True}),
    "gaussian": ({"shape": "gaussian", "size": 10.0, "ocoords": "altaz", "orientation": NP.asarray([90.0, 270.0])}, {"pc": True}),
    "dipole_gp_mod": ({"shape": "dipole", "size": 1.2, "ocoords": "dircos", "orientation": NP.asarray([[0.0, 1.0, 0.0]]),
                       "groundplane": 0.4, "ground_modify": {"scale": 0.8, "max": 1.5}}, {}),
    "paper_short": ({"id": "paper", "orientation": NP.asarray([0.0, 90.0]), "ocoords": "altaz"}, {"short_dipole_approx": True}),
    "paper_halfwave": ({"id": "paper", "orientation": NP.asarray([0.0, 90.0]), "ocoords": "altaz"}, {"half_wave_dipole_approx": True}),
}


@pytest.mark.parametrize("key", sorted(BEAM_CASES))
def test_beams_match_reference_golden(key):
    """GPU beam evaluation against the golden vectors from the reference's primary_beams.py."""
    from prisim_b200 import primary_beams as PB
    g = NP.load(os.path.join(GOLD, "beams.npz"))
    tel, opt = BEAM_CASES[key]
    kw = {k: v for k, v in opt.items() if k != "pc"}
    if opt.get("pc"):
        kw["pointing_center"] = g["pc_altaz"]
    pb = PB.primary_beam_generator(g["altaz"], g["freqs_ghz"], dict(tel), skyunits="altaz", freq_scale="GHz", **kw)
    ref = g["pb_" + key]
    assert pb.shape == ref.shape
    # the device table is fp32: 2^-24 relative to the peak of each pattern
    assert NP.abs(pb - ref).max() <= 1.5e-7 * max(1.0, NP.abs(ref).max())


def test_phased_tile_beam_matches_oracle_and_reference():
    from prisim_b200 import primary_beams as PB
    g = NP.load(os.path.join(GOLD, "beams.npz"))
    tel = {"id": "mwa", "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3,
           "element_locs": g["element_locs"]}
    for key, pinfo in (("pb_mwa_tile_delays", {"delays": g["tile_delays"]}),
                       ("pb_mwa_tile_pointing", {"pointing_center": NP.asarray([52.806, 101.31]), "pointing_coords": "altaz"})):
        pb = PB.primary_beam_generator(g["altaz"], g["freqs_ghz"], dict(tel), skyunits="altaz", pointing_info=dict(pinfo))
        pb64 = O.primary_beam_generator(g["altaz"], g["freqs_ghz"], dict(tel), skyunits="altaz", pointing_info=dict(pinfo))
        assert NP.abs(pb - pb64).max() <= 1.5e-7 * NP.abs(pb64).max()          # fp64 oracle: fp32 storage rounding only
        assert NP.abs(pb - g[key]).max() <= 2e-3 * NP.abs(g[key]).max()        # reference computes this in float32


def test_amp_table_spectrum_and_layout(eng):
    from prisim_b200 import _lib
    rng = NP.random.default_rng(7)
    nsrc0, nchan = 333, 200                                                    # ragged: 2 slabs, 11 tiles
    altaz = NP.stack((rng.uniform(-20, 90, nsrc0), rng.uniform(0, 360, nsrc0)), 1)
    freqs = 120e6 + NP.arange(nchan) * 250e3
    S, alpha, off = rng.uniform(0.1, 10, nsrc0), rng.normal(-0.8, 0.3, nsrc0), rng.uniform(0, 0.1, nsrc0)
    fref = rng.uniform(100e6, 200e6, nsrc0)
    dircos, idx = eng.sky_cull(altaz, "altaz")
    m2 = idx.cpu().numpy()
    spec = {"flux_scale": eng._f64(S, 0), "index": eng._f64(alpha, 0), "freq_ref": eng._f64(fref, 0), "flux_offset": eng._f64(off, 0)}
    beam = eng.make_beam_desc(element=_lib.BEAM_DELTA)
    amp = eng.amp_table(dircos, idx, m2.size, spec, beam, freqs)
    assert amp.numel() * 4 == _lib.load().pb200_amp_bytes(m2.size, nchan)
    dense = eng.amp_table_to_dense(amp, m2.size, nchan).cpu().numpy()
    ref = O.power_law_spectrum(S[m2], alpha[m2], fref[m2], freqs, off[m2])
    assert NP.abs(dense - ref).max() <= 1.2e-7 * NP.abs(ref).max()
    # padding rows / channels are zero
    full = amp.view(2, -1, 128)
    assert full[:, m2.size:, :].abs().max().item() == 0.0 and full[1, :, nchan - 128:].abs().max().item() == 0.0
    # tabulated spectrum + beam table + achromatic option
    tab = rng.uniform(0, 5, (nsrc0, nchan))
    pbt = rng.uniform(0, 1, (m2.size, nchan))
    amp2 = eng.amp_table(dircos, idx, m2.size, {"spectrum": eng._f64(tab, 0)}, eng.make_beam_desc(element=_lib.BEAM_TABLE), freqs,
                         pbeam=eng._f64(pbt, 0))
    assert NP.abs(eng.amp_table_to_dense(amp2, m2.size, nchan).cpu().numpy() - tab[m2] * pbt).max() <= 1.2e-7 * 5
    ach = eng.make_beam_desc(element=_lib.BEAM_AIRY, size=14.0, pointing=(0, 0, 1), achromatic=1, ref_freq_hz=150e6)
    one = torch.ones(nsrc0, dtype=torch.float64, device="cuda")
    amp3 = eng.amp_table(dircos, idx, m2.size, {"flux_scale": one, "index": torch.zeros_like(one), "freq_ref": one}, ach, freqs)
    ref3 = O.airy_disk_pattern(14.0, altaz[m2], NP.asarray([150e6]), pointing_center=NP.asarray([90.0, 270.0]), pointing_coords="altaz")
    assert NP.abs(eng.amp_table_to_dense(amp3, m2.size, nchan).cpu().numpy() - ref3).max() <= 1.5e-7


# ------------------------------------------------------------------ the phase sum
def _skyvis_case(eng, rng, nsrc, bl, freqs, pc_altaz=(90.0, 270.0), taper=False, method="auto"):
    altaz = NP.stack((NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))), rng.uniform(0, 360, nsrc)), 1)
    amp_dense = rng.uniform(0.05, 5.0, (nsrc, freqs.size)) * rng.uniform(0, 1, (nsrc, 1)) ** 4
    dircos, idx = eng.sky_cull(altaz, "altaz")
    amp = eng.dense_to_amp_table(torch.as_tensor(amp_dense).cuda())
    amp32 = eng.amp_table_to_dense(amp, nsrc, freqs.size).double().cpu().numpy()
    fw = None
    shape = None
    if taper:
        fwhm = rng.uniform(0.02, 0.8, nsrc)
        fw = eng._f64(fwhm, 0)
        shape = NP.stack((fwhm, fwhm, NP.zeros(nsrc)), 1)
    V = eng.skyvis(dircos, amp, nsrc, bl, O.altaz2dircos(pc_altaz)[0], freqs, src_fwhm_deg=fw, method=method).cpu().numpy()
    Vo = O.skyvis_snapshot(bl, altaz, amp32, freqs, NP.asarray(pc_altaz), src_shape=shape)
    return V, Vo


@pytest.mark.parametrize("nsrc,nchan,scale", [(1, 2, 1.0), (33, 33, 1.0), (700, 128, 1.0), (1500, 160, 25.0), (2100, 64, 60.0)])
def test_skyvis_recurrence_vs_oracle(eng, nsrc, nchan, scale):
    rng = NP.random.default_rng(nsrc)
    freqs = 150e6 + (NP.arange(nchan) - nchan // 2) * 97656.25
    V, Vo = _skyvis_case(eng, rng, nsrc, _hex_bl(scale), freqs)
    assert rel_err_per_baseline(V, Vo) <= TOL


def test_skyvis_off_zenith_phase_centre_and_3d_baselines(eng):
    rng = NP.random.default_rng(11)
    bl = rng.normal(0, 300.0, (77, 3))
    freqs = 185e6 + (NP.arange(96) - 48) * 40e3
    V, Vo = _skyvis_case(eng, rng, 900, bl, freqs, pc_altaz=(52.806, 101.31))
    assert rel_err_per_baseline(V, Vo) <= TOL


def test_skyvis_direct_kernel_nonuniform_channels(eng):
    rng = NP.random.default_rng(12)
    freqs = NP.sort(rng.uniform(100e6, 200e6, 70))
    V, Vo = _skyvis_case(eng, rng, 400, _hex_bl(10.0), freqs)                  # auto -> direct
    assert rel_err_per_baseline(V, Vo) <= TOL
    uni = 150e6 + NP.arange(64) * 1e5
    Vd, Vo2 = _skyvis_case(eng, NP.random.default_rng(13), 400, _hex_bl(10.0), uni, method="direct")
    Vr, _ = _skyvis_case(eng, NP.random.default_rng(13), 400, _hex_bl(10.0), uni, method="recurrence")
    assert rel_err_per_baseline(Vd, Vo2) <= TOL and rel_err_per_baseline(Vr, Vo2) <= TOL
    from prisim_b200._lib import PB200Error
    with pytest.raises(PB200Error):
        _skyvis_case(eng, rng, 10, _hex_bl(), freqs, method="recurrence")


def test_skyvis_taper_vs_oracle(eng):
    rng = NP.random.default_rng(14)
    freqs = 150e6 + (NP.arange(64) - 32) * 390625.0
    V, Vo = _skyvis_case(eng, rng, 800, _hex_bl(20.0), freqs, taper=True)
    assert rel_err_per_baseline(V, Vo) <= TOL
    Vd, Vo = _skyvis_case(eng, NP.random.default_rng(14), 800, _hex_bl(20.0), freqs, taper=True, method="direct")
    assert rel_err_per_baseline(Vd, Vo) <= TOL


def test_skyvis_analytic_kats(eng):
    from prisim_b200 import _lib
    freqs = 150e6 + (NP.arange(128) - 64) * 100e3
    bl = _hex_bl(5.0)
    one = torch.ones(2, dtype=torch.float64, device="cuda")
    spec = {"flux_scale": 3.5 * one, "index": torch.zeros_like(one), "freq_ref": one}
    # source at the phase centre: V = S for every baseline and channel
    dircos, idx = eng.sky_cull(NP.asarray([[90.0, 0.0]]), "altaz")
    amp = eng.amp_table(dircos, idx, 1, spec, eng.make_beam_desc(element=_lib.BEAM_DELTA), freqs)
    V = eng.skyvis(dircos, amp, 1, bl, (0.0, 0.0, 1.0), freqs).cpu().numpy()
    assert NP.abs(V - 3.5).max() <= 1e-6
    # below-horizon sources are culled: empty ROI -> zeros
    dircos, idx = eng.sky_cull(NP.asarray([[-5.0, 10.0]]), "altaz")
    assert idx.numel() == 0
    V0 = eng.skyvis(dircos, None, 0, bl, (0.0, 0.0, 1.0), freqs)
    assert V0.abs().max().item() == 0.0


def test_skyvis_properties_linearity_conjugate_sharding(eng):
    rng = NP.random.default_rng(21)
    nsrc, nchan = 5000, 256
    freqs = 150e6 + (NP.arange(nchan) - nchan // 2) * 97656.25
    from prisim_b200 import synthetic as S
    bl = S.array_baselines(S.hera_layout(5))[0]                                # 61 antennas, 1830 baselines
    altaz = NP.stack((NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))), rng.uniform(0, 360, nsrc)), 1)
    dircos, _ = eng.sky_cull(altaz, "altaz")
    dense = torch.rand((nsrc, nchan), device="cuda", dtype=torch.float64)
    amp = eng.dense_to_amp_table(dense)
    pc = (0.0, 0.0, 1.0)
    V = eng.skyvis(dircos, amp, nsrc, bl, pc, freqs)
    # conjugate symmetry V(-b) = V(b)*  (exact: same arithmetic with the sign of every phase flipped)
    Vm = eng.skyvis(dircos, amp, nsrc, -bl, pc, freqs)
    rms = V.abs().pow(2).mean().sqrt().item()
    assert (Vm - V.conj()).abs().max().item() <= 1e-6 * rms
    # linearity in the amplitudes (power-of-two scale is exact in fp32)
    V4 = eng.skyvis(dircos, eng.dense_to_amp_table(4.0 * dense), nsrc, bl, pc, freqs)
    assert torch.equal(V4, 4.0 * V)
    # additivity over disjoint source sets
    Va = eng.skyvis(dircos[:2016].contiguous(), eng.dense_to_amp_table(dense[:2016]), 2016, bl, pc, freqs)
    Vb = eng.skyvis(dircos[2016:].contiguous(), eng.dense_to_amp_table(dense[2016:]), nsrc - 2016, bl, pc, freqs)
    assert (Va + Vb - V).abs().max().item() <= 2e-6 * rms
    # deterministic: the same launch twice is bit-identical (stream-K head partials are added in CTA order, no atomics)
    assert torch.equal(eng.skyvis(dircos, amp, nsrc, bl, pc, freqs), V)
    # baseline sharding: every (b,f) is still one owner's sum, but a different baseline count moves the stream-K split
    # points along the source axis, i.e. the grouping of the fp32 partial sums -- equal to fp32 rounding, not bit-exact
    for parts in (2, 3, 8):
        from prisim_b200.sharding import shard_bounds
        b = shard_bounds(bl.shape[0], parts)
        Vs = torch.cat([eng.skyvis(dircos, amp, nsrc, bl[b[r]:b[r + 1]], pc, freqs) for r in range(parts)], dim=0)
        assert (Vs - V).abs().max().item() <= 2e-6 * rms
    # the fp64 kernel reproduces itself across shardings to fp64 rounding
    amp64 = eng.dense_to_amp_table(dense, dtype=torch.float64)
    V64 = eng.skyvis(dircos, amp64, nsrc, bl, pc, freqs, method="fp64")
    b = shard_bounds(bl.shape[0], 3)
    V64s = torch.cat([eng.skyvis(dircos, amp64, nsrc, bl[b[r]:b[r + 1]], pc, freqs, method="fp64") for r in range(3)], dim=0)
    assert (V64s - V64).abs().max().item() <= 1e-12 * rms
    assert (V - V64).abs().max().item() <= 1e-5 * rms


@pytest.mark.parametrize("nbl,nchan,nsrc", [(64 * 150 + 5, 256, 96), (128 * 149, 128, 64), (64 * 297, 300, 40)])
def test_skyvis_persistent_schedule_waves_and_tail(eng, nbl, nchan, nsrc):
    """More output tiles than SMs: whole-tile waves plus a stream-K tail split along the source axis (head partials added
    by k_skyvis_finalize).  fp32 recurrence, direct and fp64 kernels against each other on every cell and against the
    oracle on sampled rows."""
    rng = NP.random.default_rng(nbl)
    freqs = 150e6 + (NP.arange(nchan) - nchan // 2) * 97656.25
    bl = rng.normal(0.0, 150.0, (nbl, 3)) * NP.asarray([1.0, 1.0, 0.02])
    altaz = NP.stack((NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))), rng.uniform(0, 360, nsrc)), 1)
    dense = rng.uniform(0.05, 5.0, (nsrc, nchan)) * rng.uniform(0, 1, (nsrc, 1)) ** 4
    dircos, _ = eng.sky_cull(altaz, "altaz")
    amp = eng.dense_to_amp_table(torch.as_tensor(dense).cuda())
    amp64 = eng.dense_to_amp_table(eng.amp_table_to_dense(amp, nsrc, nchan).double(), dtype=torch.float64)   # the fp32-rounded amplitudes
    pc = (0.0, 0.0, 1.0)
    V64 = eng.skyvis(dircos, amp64, nsrc, bl, pc, freqs, method="fp64")
    rms_b = V64.abs().pow(2).mean(dim=1, keepdim=True).sqrt()
    for method in ("recurrence", "direct") + (("recurrence_pair",) if nchan % 256 == 0 else ()):
        V = eng.skyvis(dircos, amp, nsrc, bl, pc, freqs, method=method)
        assert ((V - V64).abs() / rms_b).max().item() <= TOL, method
    if nchan % 256:                                    # the pair form (two baselines x 16 channels per thread) says so when it does not apply
        with pytest.raises(Exception, match="pair form"):
            eng.skyvis(dircos, amp, nsrc, bl, pc, freqs, method="recurrence_pair")
    rows = NP.unique(NP.concatenate(([0, nbl - 1], rng.choice(nbl, 40, replace=False))))
    amp32 = eng.amp_table_to_dense(amp, nsrc, nchan).double().cpu().numpy()
    Vo = O.skyvis_snapshot(bl[rows], altaz, amp32, freqs, NP.asarray([90.0, 270.0]))
    got = V64[torch.as_tensor(rows).cuda()].cpu().numpy()
    assert float((NP.abs(got - Vo) / rms_b[torch.as_tensor(rows).cuda()].cpu().numpy()).max()) <= 1e-10


def test_skyvis_full_size_config2_subset_parity(eng):
    """BASELINE config 2 at full size on the GPU (61,075 baselines x 1024 channels x 300k-source
    catalogue, ~179k above the horizon); parity on a baseline/channel subset the oracle can do."""
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    cfg = S.config2()
    ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", pointing_coords="hadec", device=0)
    ia.observe(SimpleTime(2451545.0, 0.0), {"Tnet": 300.0}, NP.ones(1024), cfg["pointing_hadec"], cfg["skymodel"], cfg["t_acc"])
    V = ia.skyvis_freq_device(0)
    m2 = ia.obs_catalog_indices[0]
    assert 170000 < m2.size < 190000                 # Dec < +30 deg catalogue seen from lat -30.7
    rng = NP.random.default_rng(2)
    bsel = NP.sort(NP.concatenate(([0, 61074], rng.choice(61075, 14, replace=False))))
    csel = NP.sort(rng.choice(1024, 48, replace=False))
    sky = cfg["skymodel"]
    hadec = NP.stack((0.0 - sky.location[:, 0], sky.location[:, 1]), 1)
    sp = sky.spec_parms
    Vo, m2o = O.observe_snapshot(cfg["baselines"][bsel], cfg["channels"][csel], hadec, "hadec", cfg["latitude"], cfg["pointing_hadec"],
                                 "hadec", cfg["telescope"], sp["flux-scale"], sp["power-law-index"], sp["freq-ref"])
    assert NP.array_equal(m2, m2o)
    Vg = V[torch.as_tensor(bsel).cuda()][:, torch.as_tensor(csel).cuda()].cpu().numpy()
    rms_b = V[torch.as_tensor(bsel).cuda()].abs().pow(2).mean(dim=1, keepdim=True).sqrt().cpu().numpy()
    assert float((NP.abs(Vg - Vo) / rms_b).max()) <= TOL


def test_skyvis_full_grid_config2_fp32_vs_fp64_every_cell(eng):
    """Config 2 at full size: the fp32 kernel (the one bench.py times) against the fp64 kernel (oracle-validated to 1e-10
    above) on EVERY one of the 61,075 x 1024 cells: max |dV| <= 1e-5 rms_b.  The oracle itself would need core-weeks."""
    from prisim_b200 import primary_beams as PB
    from prisim_b200 import synthetic as S
    cfg = S.config2()
    sky, sp = cfg["skymodel"], cfg["skymodel"].spec_parms
    hadec = eng._f64(NP.stack((0.0 - sky.location[:, 0], sky.location[:, 1]), axis=1), 0)
    spec = {"flux_scale": eng._f64(sp["flux-scale"], 0), "index": eng._f64(sp["power-law-index"], 0), "freq_ref": eng._f64(sp["freq-ref"], 0)}
    beam = PB.beam_desc_from_telescope(cfg["telescope"], pointing_center=NP.asarray([90.0, 270.0]), skyunits="altaz", device=0)
    dircos, index = eng.sky_cull(hadec, "hadec", latitude_deg=cfg["latitude"])
    nsrc = int(index.shape[0])
    bl = eng._f64(cfg["baselines"], 0)
    amp = eng.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"])
    V = eng.skyvis(dircos, amp, nsrc, bl, (0.0, 0.0, 1.0), cfg["channels"])
    del amp
    amp64 = eng.amp_table(dircos, index, nsrc, spec, beam, cfg["channels"], dtype=torch.float64)
    V64 = eng.skyvis(dircos, amp64, nsrc, bl, (0.0, 0.0, 1.0), cfg["channels"], method="fp64")
    del amp64
    rms_b = V64.abs().pow(2).mean(dim=1, keepdim=True).sqrt()
    err = ((V - V64).abs() / rms_b).amax(dim=1)
    worst = float(err.max().item())
    print("config 2, all {0} x 1024 cells: max |dV|/rms_b = {1:.3e} (median over baselines {2:.3e})".format(bl.shape[0], worst, float(err.median().item())))
    assert worst <= TOL


# ------------------------------------------------------------------ noise
def test_module_level_generate_noise(eng):
    """generateNoise (interferometry.py:236-329): rms/sqrt2 (N + iN) of shape (nbl, nchan, ntimes) from given rms or from
    the instrument parameters; device generator, reproducible by seed."""
    from prisim_b200.interferometry import generateNoise, thermalNoiseRMS
    nbl, nchan, nt = 64, 96, 3
    rms = NP.linspace(1.0, 3.0, nchan).reshape(1, nchan, 1)
    nz = generateNoise(noiseRMS=rms, nbl=nbl, nchan=nchan, ntimes=nt, seed=5, device=0)
    assert nz.shape == (nbl, nchan, nt) and nz.dtype == NP.complex128
    z = nz / (rms / NP.sqrt(2.0))
    n = z.size
    assert abs(z.real.std() - 1) < 5 / NP.sqrt(2 * n) and abs(z.imag.std() - 1) < 5 / NP.sqrt(2 * n) and abs(z.mean()) < 5 / NP.sqrt(n)
    assert abs(NP.mean(z.real[..., 0] * z.real[..., 1])) < 5 / NP.sqrt(n / nt)           # time slices are independent
    assert NP.array_equal(nz, generateNoise(noiseRMS=rms, nbl=nbl, nchan=nchan, ntimes=nt, seed=5, device=0))
    assert not NP.array_equal(nz, generateNoise(noiseRMS=rms, nbl=nbl, nchan=nchan, ntimes=nt, seed=6, device=0))
    nz2 = generateNoise(A_eff=100.0, df=1e5, dt=10.0, Tsys=200.0, nbl=nbl, nchan=nchan, ntimes=1, eff_Q=0.9, seed=1, device=0)
    r0 = float(thermalNoiseRMS(100.0, 1e5, 10.0, 200.0, eff_Q=0.9)[0, 0, 0])
    assert abs(NP.sqrt(NP.mean(NP.abs(nz2) ** 2)) / r0 - 1) < 5 / NP.sqrt(nz2.size)
    with pytest.raises(IndexError):
        generateNoise(noiseRMS=NP.ones((2, 5)), nbl=2, nchan=5, ntimes=1)


def test_noise_rms_statistics_and_sharding_invariance(eng):
    nbl, nchan = 300, 128
    freqs = 150e6 + NP.arange(nchan) * 1e5
    Tsys = 50.0 + 200.0 * (freqs / 150e6) ** -2.55
    tsys = eng._f64(Tsys, 0)
    aeff = eng._f64(NP.full((nbl, nchan), 100.1), 0)
    effq = eng._f64([0.96], 0)
    sky = torch.complex(torch.rand((nbl, nchan), dtype=torch.float64, device="cuda"), torch.rand((nbl, nchan), dtype=torch.float64, device="cuda"))
    rms, nz, vis = eng.noise(sky, tsys, aeff, effq, 1e5, 10.7, seed=99, nbl=nbl, nchan=nchan, snapshot=3)
    ref = O.thermal_noise_rms(NP.repeat(Tsys[None, :, None], nbl, axis=0), 100.1, 0.96, [10.7], 1e5)[:, :, 0]
    assert NP.abs(rms.cpu().numpy() - ref).max() <= 1e-13 * ref.max()
    z = (nz / (rms / NP.sqrt(2.0))).cpu().numpy()                               # should be N(0,1) + i N(0,1)
    n = z.size
    assert abs(z.real.mean()) < 5 / NP.sqrt(n) and abs(z.imag.mean()) < 5 / NP.sqrt(n)
    assert abs(z.real.std() - 1) < 5 / NP.sqrt(2 * n) and abs(z.imag.std() - 1) < 5 / NP.sqrt(2 * n)
    assert abs(NP.mean(z.real * z.imag)) < 5 / NP.sqrt(n)
    assert abs(NP.mean(z.real[:, 1:] * z.real[:, :-1])) < 5 / NP.sqrt(n)        # white along frequency
    assert abs(NP.mean(z.real ** 4) - 3.0) < 0.1                               # Gaussian kurtosis
    assert torch.equal(vis, sky + nz)
    assert torch.equal(eng.add_noise(sky, nz), vis)
    # same seed -> same numbers; different snapshot / seed -> different
    _, nz2, _ = eng.noise(None, tsys, aeff, effq, 1e5, 10.7, seed=99, nbl=nbl, nchan=nchan, snapshot=3, want=("noise",))
    _, nz3, _ = eng.noise(None, tsys, aeff, effq, 1e5, 10.7, seed=99, nbl=nbl, nchan=nchan, snapshot=4, want=("noise",))
    assert torch.equal(nz, nz2) and not torch.equal(nz, nz3)
    # sharding invariance: rows [100:250) generated as a shard equal the same rows of the full run
    _, nzs, _ = eng.noise(None, tsys, aeff[100:250].contiguous(), effq, 1e5, 10.7, seed=99, nbl=150, nchan=nchan, snapshot=3,
                          bl_offset=100, nbl_total=nbl, want=("noise",))
    assert torch.equal(nzs, nz[100:250])
    # odd channel count and odd shard offsets: Philox word pairs straddle rows and shard boundaries
    nco = 127
    tso, aeo = eng._f64(Tsys[:nco], 0), eng._f64(NP.full((nbl, nco), 100.1), 0)
    _, nzo, _ = eng.noise(None, tso, aeo, effq, 1e5, 10.7, seed=7, nbl=nbl, nchan=nco, snapshot=1, want=("noise",))
    for lo, hi in ((0, 1), (1, 2), (37, 154), (299, 300)):
        _, part, _ = eng.noise(None, tso, aeo[lo:hi].contiguous(), effq, 1e5, 10.7, seed=7, nbl=hi - lo, nchan=nco, snapshot=1,
                               bl_offset=lo, nbl_total=nbl, want=("noise",))
        assert torch.equal(part, nzo[lo:hi])
    zo = (nzo / nzo.abs().pow(2).mean().sqrt()).cpu().numpy()
    assert abs(NP.mean(zo.real[:, 1:] * zo.real[:, :-1])) < 5 / NP.sqrt(zo.size) and abs(NP.mean(zo.real * zo.imag)) < 5 / NP.sqrt(zo.size)
    # K units (interferometry.py:6689)
    rmsk, _, _ = eng.noise(None, tsys, None, effq, 1e5, 10.7, seed=1, nbl=nbl, nchan=nchan, flux_unit_k=True, want=("rms",))
    assert NP.abs(rmsk.cpu().numpy() - Tsys[None, :] / 0.96 / NP.sqrt(10.7e5)).max() < 1e-12


# ------------------------------------------------------------------ delay transform
@pytest.mark.parametrize("nchan,pad", [(128, 1.0), (128, 0.0), (128, 0.5), (128, 2.0), (1024, 1.0), (100, 1.0), (100, 0.0),
                                       (96, 0.5), (33, 1.0), (33, 0.0), (2, 1.0), (4096, 1.0)])
def test_delay_transform_vs_oracle(eng, nchan, pad):
    rng = NP.random.default_rng(nchan)
    nrows = 37
    x = rng.standard_normal((nrows, nchan)) + 1j * rng.standard_normal((nrows, nchan))
    bp = 1.0 + 0.1 * rng.standard_normal((nrows, nchan))
    w = O.windowing(nchan, "bhw", area_normalize=True) * nchan if nchan > 2 else NP.ones(nchan)
    df = 97656.25
    out = eng.delay_transform(torch.as_tensor(x).cuda(), eng._f64(bp, 0), eng._f64(w, 0), df, pad=pad).cpu().numpy()
    ref, lags = O.delay_transform(x[:, :, None], bp[:, :, None], NP.repeat(w[None, :, None], nrows, axis=0), df, pad=pad)
    ref = ref[:, :, 0]
    assert out.shape == ref.shape
    assert rel_err_per_baseline(out, ref) <= 1e-11
    # no downsampling, lag kernel (x = None) and Parseval on the unpadded transform
    out2 = eng.delay_transform(torch.as_tensor(x).cuda(), None, None, df, pad=pad, downsample=False).cpu().numpy()
    ref2, _ = O.delay_transform(x[:, :, None], NP.ones((nrows, nchan, 1)), NP.ones((nrows, nchan, 1)), df, pad=pad, downsample=False)
    assert rel_err_per_baseline(out2, ref2[:, :, 0]) <= 1e-11
    kern = eng.delay_transform(None, eng._f64(bp, 0), eng._f64(w, 0), df, pad=pad, nrows=nrows, nchan=nchan, device=0).cpu().numpy()
    refk, _ = O.delay_transform(NP.ones((nrows, nchan, 1)), bp[:, :, None], NP.repeat(w[None, :, None], nrows, axis=0), df, pad=pad)
    assert rel_err_per_baseline(kern, refk[:, :, 0]) <= 1e-11
    if pad == 0.0:
        assert NP.allclose(NP.sum(NP.abs(out2) ** 2, axis=1) / (nchan * df ** 2), NP.sum(NP.abs(x) ** 2, axis=1), rtol=1e-10)


def test_skyvis_random_shapes_property(eng):
    """Property test over ragged shapes (hypothesis): any (nsrc, nbl, nchan) -- source counts that are not a
    multiple of the 32-row tile, baseline counts that leave partially filled CTAs, channel counts that leave
    partially filled or missing slabs -- matches the oracle, for every kernel variant."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=14, deadline=None, suppress_health_check=list(HealthCheck))
    @given(nsrc=st.integers(1, 200), nbl=st.integers(1, 150), nchan=st.integers(2, 300), seed=st.integers(0, 10 ** 6),
           method=st.sampled_from(["auto", "recurrence_lift", "recurrence_3term", "recurrence_3term_scalar", "recurrence_quarter", "recurrence_scalar", "direct", "fp64"]), spc=st.sampled_from(["1", "2", "4"]))
    def check(nsrc, nbl, nchan, seed, method, spc):
        rng = NP.random.default_rng(seed)
        bl = rng.normal(0, 80.0, (nbl, 3)); bl[:, 2] *= 0.05
        freqs = 150e6 + (NP.arange(nchan) - nchan // 2) * 97656.25
        altaz = NP.stack((NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))), rng.uniform(0, 360, nsrc)), 1)
        dense = rng.uniform(0.05, 5.0, (nsrc, nchan))
        dircos, _ = eng.sky_cull(altaz, "altaz")
        amp = eng.dense_to_amp_table(torch.as_tensor(dense).cuda(), dtype=torch.float64 if method == "fp64" else torch.float32)
        amp_used = eng.amp_table_to_dense(amp, nsrc, nchan).double().cpu().numpy()
        from prisim_b200 import _lib
        ctx = _lib.get_context(0)
        ctx.set_option("skyvis_spc", int(spc))
        try:
            V = eng.skyvis(dircos, amp, nsrc, bl, (0.0, 0.0, 1.0), freqs, method=method).cpu().numpy()
        finally:
            ctx.set_option("skyvis_spc", 0)
        Vo = O.skyvis_snapshot(bl, altaz, amp_used, freqs, NP.asarray([90.0, 270.0]))
        assert V.shape == (nbl, nchan)
        assert rel_err_per_baseline(V, Vo) <= (1e-11 if method == "fp64" else TOL)

    check()
