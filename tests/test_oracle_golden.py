"""Oracle vs golden vectors produced by executing the reference's own source
(tests/golden/make_golden.py).  Runs on CPU; this is what pins the oracle."""
import os

import numpy as NP
import pytest

from oracle import prisim_oracle as O

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

TELESCOPES = {
    "hera": {"id": "hera", "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz"},
    "hirax_offzenith": {"id": "hirax", "orientation": NP.asarray([75.0, 120.0]), "ocoords": "altaz"},
    "mwa_dipole_gp": {"id": "mwa_dipole", "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3},
    "paper": {"id": "paper", "orientation": NP.asarray([0.0, 90.0]), "ocoords": "altaz"},
    "mwa_analytic": {"id": "mwa", "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3},
    "delta": {"shape": "delta"},
    "dish": {"shape": "dish", "size": 14.0, "ocoords": "altaz", "orientation": NP.asarray([90.0, 270.0])},
    "gaussian": {"shape": "gaussian", "size": 10.0, "ocoords": "altaz", "orientation": NP.asarray([90.0, 270.0])},
    "dipole_gp_mod": {"shape": "dipole", "size": 1.2, "ocoords": "dircos", "orientation": NP.asarray([[0.0, 1.0, 0.0]]),
                      "groundplane": 0.4, "ground_modify": {"scale": 0.8, "max": 1.5}},
}

OBSERVE_CASES = {
    "hera": dict(telescope={"id": "hera", "shape": "dish", "size": 14.0, "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz", "groundplane": None}),
    "hera_taper": dict(telescope={"id": "hera", "shape": "dish", "size": 14.0, "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz", "groundplane": None}),
    "hera_taper_gradient": dict(telescope={"id": "hera", "shape": "dish", "size": 14.0, "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz", "groundplane": None}),
    "hera_roi20": dict(telescope={"id": "hera", "shape": "dish", "size": 14.0, "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz", "groundplane": None}, roi_radius=20.0),
    "hera_roiinfo": dict(telescope={"id": "hera", "shape": "dish", "size": 14.0, "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz", "groundplane": None}),
    "gaussian_altazpointing": dict(telescope={"shape": "gaussian", "size": 10.0, "ocoords": "altaz", "orientation": NP.asarray([90.0, 270.0]), "groundplane": None},
                                   pointing_coords="altaz"),
    "mwa_dipole": dict(telescope={"id": "mwa_dipole", "shape": "dipole", "size": 0.74, "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3}),
}


def _load(name):
    return NP.load(os.path.join(GOLD, name), allow_pickle=False)


@pytest.mark.parametrize("key", sorted(TELESCOPES))
def test_beam_presets_match_reference(key):
    g = _load("beams.npz")
    kw = dict(skyunits="altaz", freq_scale="GHz")
    if key in ("dish", "gaussian"):
        kw["pointing_center"] = g["pc_altaz"]
    pb = O.primary_beam_generator(g["altaz"], g["freqs_ghz"], dict(TELESCOPES[key]), **kw)
    ref = g["pb_" + key]
    assert pb.shape == ref.shape
    assert NP.allclose(pb, ref, rtol=1e-9, atol=1e-12)


def test_dipole_approximations_match_reference():
    g = _load("beams.npz")
    t = TELESCOPES["paper"]
    assert NP.allclose(O.primary_beam_generator(g["altaz"], g["freqs_ghz"], dict(t), skyunits="altaz", short_dipole_approx=True),
                       g["pb_paper_short"], rtol=1e-9, atol=1e-12)
    assert NP.allclose(O.primary_beam_generator(g["altaz"], g["freqs_ghz"], dict(t), skyunits="altaz", half_wave_dipole_approx=True),
                       g["pb_paper_halfwave"], rtol=1e-9, atol=1e-12)


def test_phased_tile_matches_reference_in_float32_and_float64():
    """The reference evaluates the phased-array factor in float32 (primary_beams.py:1730-1746).
    reference_float32=True reproduces it to float32 rounding; the float64 evaluation (the parity
    target of the GPU kernel) deviates from it only at the float32 level."""
    g = _load("beams.npz")
    tel = {"id": "mwa", "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3,
           "element_locs": g["element_locs"]}
    for key, pinfo in (("pb_mwa_tile_delays", {"delays": g["tile_delays"]}),
                       ("pb_mwa_tile_pointing", {"pointing_center": NP.asarray([52.806, 101.31]), "pointing_coords": "altaz"})):
        ref = g[key]
        pb32 = O.primary_beam_generator(g["altaz"], g["freqs_ghz"], dict(tel), skyunits="altaz", pointing_info=dict(pinfo),
                                        reference_float32=True)
        pb64 = O.primary_beam_generator(g["altaz"], g["freqs_ghz"], dict(tel), skyunits="altaz", pointing_info=dict(pinfo))
        assert NP.abs(pb32 - ref).max() <= 2e-6 * NP.abs(ref).max()
        assert NP.abs(pb64 - ref).max() <= 2e-3 * NP.abs(ref).max()


def test_geometric_delay_matches_reference():
    g = _load("delays.npz")
    lat = float(g["latitude"])
    assert NP.allclose(O.geometric_delay(g["bl"], g["altaz"], altaz=True, hadec=False), g["tau_altaz"], rtol=0, atol=1e-20)
    assert NP.allclose(O.geometric_delay(g["bl"], g["hadec"], altaz=False, hadec=True, latitude=lat), g["tau_hadec"], rtol=0, atol=1e-20)
    assert NP.allclose(O.geometric_delay(g["bl"], O.altaz2dircos(g["altaz"]), altaz=False, hadec=False, dircos=True), g["tau_dircos"], rtol=0, atol=1e-20)
    hz = O.horizon_delay_limits(g["bl"], O.altaz2dircos([90.0, 270.0])[0])
    assert NP.allclose(hz, g["horizon"].reshape(hz.shape), rtol=1e-12, atol=1e-20)


def oracle_run(g, case):
    """Re-run a golden observe() case through the oracle; returns dict of products."""
    nsnap = int(g["n_acc"])
    lat = float(g["latitude"])
    src_shape = g["src_shape"] if g["src_shape"].size else None
    pointing_coords = case.get("pointing_coords", "hadec")
    skyvis, m2s = [], []
    for j in range(nsnap):
        roi_info = None
        if "roi_ind_{0}".format(j) in g.files:
            roi_info = {"ind": g["roi_ind_{0}".format(j)], "pbeam": g["roi_pbeam_{0}".format(j)]}
        V, m2 = O.observe_snapshot(g["bl"], g["chans"], g["hadec_{0}".format(j)], "hadec", lat, g["pointing"], pointing_coords,
                                   dict(case["telescope"]), g["flux"], g["spindex"], 150e6, src_shape=src_shape,
                                   roi_radius=case.get("roi_radius", None), roi_info=roi_info, lst=float(g["lsts"][j]))
        skyvis.append(V)
        m2s.append(m2)
    return NP.stack(skyvis, axis=2), m2s


@pytest.mark.parametrize("tag", sorted(OBSERVE_CASES))
def test_observe_matches_reference(tag):
    g = _load("observe_{0}.npz".format(tag))
    case = OBSERVE_CASES[tag]
    skyvis, m2s = oracle_run(g, case)
    ref = g["skyvis_freq"]
    assert skyvis.shape == ref.shape
    scale = NP.sqrt(NP.mean(NP.abs(ref) ** 2))
    assert NP.abs(skyvis - ref).max() <= 1e-11 * scale
    for j, m2 in enumerate(m2s):
        assert NP.array_equal(m2, g["m2_{0}".format(j)])
    nbl, nchan, nsnap = ref.shape
    # Tsys, rms, noise (same legacy-RNG draws), vis
    Tsysinfo = {"Trx": 50.0, "Tant": {"T0": 200.0, "f0": 150e6, "spindex": -2.55}, "Tnet": None}
    Tsys = NP.repeat(O.system_temperature(Tsysinfo, g["chans"], nbl)[:, :, None], nsnap, axis=2)
    assert NP.allclose(Tsys, g["Tsys"])
    rms = O.thermal_noise_rms(Tsys, 154.0 * 0.65, 0.96, g["t_acc"], g["chans"][1] - g["chans"][0])
    assert NP.allclose(rms, g["vis_rms_freq"], rtol=1e-13)
    NP.random.seed(int(g["noise_seed"]))
    nre = NP.random.randn(nbl, nchan, nsnap)
    nim = NP.random.randn(nbl, nchan, nsnap)
    noise = O.noise_from_normals(rms, nre, nim)
    assert NP.allclose(noise, g["vis_noise_freq"], rtol=1e-13)
    assert NP.allclose(O.add_noise(ref, noise), g["vis_freq"], rtol=1e-13)
    # delay transforms: pad 1.0 (default), 0.0 and 0.5
    bp = g["bp"]
    wts = O.broadcast_freq_wts(g["window"], nbl, nchan, nsnap)
    df = g["chans"][1] - g["chans"][0]
    for pad, key in ((1.0, "skyvis_lag"), (0.0, "skyvis_lag_pad0"), (0.5, "skyvis_lag_pad05")):
        lag, lags = O.delay_transform(ref, bp, wts, df, pad=pad)
        assert lag.shape == g[key].shape
        assert NP.abs(lag - g[key]).max() <= 1e-12 * NP.abs(g[key]).max()
    lag, lags = O.delay_transform(g["vis_freq"], bp, wts, df, pad=1.0)
    assert NP.abs(lag - g["vis_lag"]).max() <= 1e-12 * NP.abs(g["vis_lag"]).max()
    assert NP.allclose(lags, g["lags"])
    kern, _ = O.delay_transform(NP.ones_like(ref), bp, wts, df, pad=1.0)
    assert NP.abs(kern - g["lag_kernel"]).max() <= 1e-12 * NP.abs(g["lag_kernel"]).max()


def test_rotate_visibilities_matches_reference():
    """phase_centering + project_baselines (interferometry.py:7655-7995) replayed through the oracle."""
    g = _load("observe_hera.npz")
    r = _load("rotate_hera.npz")
    lat = float(g["latitude"])
    lst = g["lsts"]
    nsnap = lst.size
    cur = O.altaz2dircos(O.hadec2altaz(NP.repeat(g["pointing"].reshape(1, 2), nsnap, axis=0), lat))
    new1_hadec = NP.repeat(r["ref1"], nsnap, axis=0)
    new1 = O.altaz2dircos(O.hadec2altaz(new1_hadec, lat))
    V1 = O.phase_rotate(g["skyvis_freq"], g["bl"], cur, new1, g["chans"])
    scale = NP.sqrt(NP.mean(NP.abs(r["skyvis_rot1"]) ** 2))
    assert NP.abs(V1 - r["skyvis_rot1"]).max() <= 1e-11 * scale
    assert NP.abs(O.phase_rotate(g["vis_freq"], g["bl"], cur, new1, g["chans"]) - r["vis_rot1"]).max() <= 1e-11 * scale
    assert NP.allclose(r["pc_rot1"], new1_hadec)
    assert NP.allclose(O.project_baselines(g["bl"], new1_hadec[:, 0], new1_hadec[:, 1], lat), r["proj_rot1"], atol=1e-10)
    new2_hadec = NP.stack((lst - r["ref2"][:, 0], r["ref2"][:, 1]), axis=1)
    new2 = O.altaz2dircos(O.hadec2altaz(new2_hadec, lat))
    V2 = O.phase_rotate(V1, g["bl"], new1, new2, g["chans"])
    assert NP.abs(V2 - r["skyvis_rot2"]).max() <= 1e-11 * scale
    assert NP.allclose(r["pc_rot2"], new2_hadec) and str(r["pc_coords"]) == "hadec"
    assert NP.allclose(O.project_baselines(g["bl"], new2_hadec[:, 0], new2_hadec[:, 1], lat), r["proj_rot2"], atol=1e-10)
    # rotating to the new centre equals simulating with that phase centre in the first place
    V_direct, _ = O.observe_snapshot(g["bl"], g["chans"], g["hadec_0"], "hadec", lat, g["pointing"], "hadec",
                                     dict(OBSERVE_CASES["hera"]["telescope"]), g["flux"], g["spindex"], 150e6)
    pbf_free = V1[:, :, 0]       # same beam (pointing unchanged), phases referred to new1
    s_new = new1[0] - cur[0]
    expect = V_direct * NP.exp(+2j * NP.pi * (g["bl"] @ s_new)[:, None] * g["chans"][None, :] / 299792458.0)
    assert NP.abs(pbf_free - expect).max() <= 1e-9 * scale


def test_gradient_matches_reference():
    """gradient_mode='baseline' (interferometry.py:6312-6343, :6384-6394) and apply_gradients (:6726-6819) executed by the
    reference's own code (tapered sky: the only case in which the reference's gradient branch runs, :6263 vs :6343)."""
    g = _load("observe_hera_taper_gradient.npz")
    case = OBSERVE_CASES["hera_taper_gradient"]
    lat = float(g["latitude"])
    grads = []
    for j in range(int(g["n_acc"])):
        (V, G), m2 = O.observe_snapshot(g["bl"], g["chans"], g["hadec_{0}".format(j)], "hadec", lat, g["pointing"], "hadec",
                                        dict(case["telescope"]), g["flux"], g["spindex"], 150e6, src_shape=g["src_shape"],
                                        lst=float(g["lsts"][j]), gradient=True)
        assert NP.abs(V - g["skyvis_freq"][:, :, j]).max() <= 1e-11 * NP.sqrt(NP.mean(NP.abs(g["skyvis_freq"]) ** 2))
        grads.append(G)
    grad = NP.stack(grads, axis=3)
    ref = g["gradient_baseline"]
    assert grad.shape == ref.shape
    assert NP.abs(grad - ref).max() <= 1e-11 * NP.sqrt(NP.mean(NP.abs(ref) ** 2))
    dV = O.apply_gradients(ref, g["perturbations"], g["chans"])
    assert dV.shape == g["delta_skyvis_freq"].shape
    assert NP.abs(dV - g["delta_skyvis_freq"]).max() <= 1e-12 * NP.abs(g["delta_skyvis_freq"]).max()
    # first-order consistency: V(b + db) - V(b) = dV + O(db^2) on the oracle itself
    j = 0
    db = g["perturbations"][0].T * 0.05                                  # [nbl,3], ~1 mm
    kw = dict(src_shape=g["src_shape"], lst=float(g["lsts"][j]))
    V1, _ = O.observe_snapshot(g["bl"] + db, g["chans"], g["hadec_0"], "hadec", lat, g["pointing"], "hadec", dict(case["telescope"]),
                               g["flux"], g["spindex"], 150e6, **kw)
    V0 = g["skyvis_freq"][:, :, 0]
    lin = O.apply_gradients(ref[:, :, :, :1], db.T[NP.newaxis], g["chans"])[0, :, :, 0]
    # the gradient covers the fringe term exp(-2 pi i f b.s/c) only; observe() also re-references the phase to the
    # (zenith) phase centre with the perturbed baseline, d/db of exp(+2 pi i f b.s_pc/c), and the taper moves slightly
    wl = 299792458.0 / g["chans"]
    full = lin + 1j * 2.0 * NP.pi / wl[NP.newaxis, :] * db[:, 2:3] * V0
    assert NP.abs((V1 - V0) - full).max() <= 0.02 * NP.abs(full).max()


def _dup_groups(g):
    keys = [tuple(k) for k in g["group_keys"].tolist()]
    return {keys[i]: [tuple(l) for l in g["group_{0}".format(i)].tolist()] for i in range(len(keys))}


def test_duplicate_measurements_index_logic_matches_reference():
    """interferometry.py:6823-6907 executed by the reference: label expansion and row repeats."""
    g = _load("duplicate.npz")
    labels = [tuple(l) for l in g["ulabels"].tolist()]
    num, out = O.duplicate_counts(labels, _dup_groups(g))
    assert out == [tuple(l) for l in g["labels_out"].tolist()]
    assert NP.array_equal(NP.repeat(g["bl"], num, axis=0), g["baselines_out"])
    assert NP.array_equal(NP.repeat(g["skyvis_unique"], num, axis=0), g["skyvis_out"])
    assert NP.array_equal(NP.repeat(g["projected_unique"], num, axis=0), g["projected_out"])
    assert NP.allclose(g["vis_minus_noise"], g["skyvis_out"])                       # noise regenerated on the expanded set
    assert tuple(g["vis_noise_shape"]) == g["skyvis_out"].shape
    # a label in two groups is an error; a key absent from labels (either order) too; complete groups are a no-op
    bad = _dup_groups(g); bad[("3", "0")] = [("3", "0"), ("2", "1")]
    with pytest.raises(ValueError):
        O.duplicate_counts(labels, bad)
    with pytest.raises(KeyError):
        O.duplicate_counts(labels, {("9", "8"): [("9", "8"), ("7", "6"), ("5", "4"), ("3", "2"), ("1", "0")]})
    assert O.duplicate_counts(labels, {("1", "0"): [("1", "0")]})[0] is None


def test_multi_window_delay_transform_matches_reference():
    """interferometry.py:8141-8287 run by the reference (window helpers of astroutils stubbed, [AU-memory])."""
    g = _load("observe_hera.npz")
    m = _load("multiwin_hera.npz")
    df = g["chans"][1] - g["chans"][0]
    for shape, bw, fc, sfx, pad in (("bhw", m["bw_eff"], m["freq_center"], "pad1", 1.0), ("bhw", m["bw_eff"], m["freq_center"], "pad0", 0.0),
                                    ("rect", 1.2e6, None, "rect", 1.0)):
        wts = O.multi_window_weights(g["chans"], bw, fc, shape)
        lag, corr = O.multi_window_delay_transform(g["skyvis_freq"], g["bp"], wts, df, pad=pad)
        ref = m["skyvis_lag_" + sfx]
        assert lag.shape == ref.shape
        assert NP.abs(lag - ref).max() <= 1e-12 * NP.abs(ref).max()
        assert NP.allclose(corr, m["lag_corr_length_" + sfx], rtol=1e-12)
        if sfx != "rect":
            nz, _ = O.multi_window_delay_transform(m["vis_noise_freq"], g["bp"], wts, df, pad=pad)
            assert NP.abs(nz - m["vis_noise_lag_" + sfx]).max() <= 1e-12 * NP.abs(m["vis_noise_lag_" + sfx]).max()
            kern, _ = O.multi_window_delay_transform(NP.ones_like(g["bp"]), g["bp"], wts, df, pad=pad)
            assert NP.abs(kern - m["lag_kernel_" + sfx]).max() <= 1e-12 * NP.abs(m["lag_kernel_" + sfx]).max()
    assert abs(O.window_N2width("bhw") - 0.35875) < 1e-5 and O.window_N2width("rect") == 1.0
    with pytest.raises(ValueError):
        O.multi_window_weights(g["chans"], [1e6, 2e6], [150e6, 150.5e6, 151e6])
    with pytest.raises(ValueError):
        O.multi_window_weights(g["chans"], 1e6, g["chans"][0])


def test_subband_delay_transform_matches_reference():
    """delay_spectrum.py:1842-2248 ('sim' branch) run by the reference's own DelaySpectrum (astroutils window helpers and the
    FFT downsampler stubbed, [AU-memory])."""
    m = _load("subband_hera.npz")
    df = m["chans"][1] - m["chans"][0]
    wts = O.subband_weights(m["chans"], m["bw_eff"], m["freq_center"], "bhw")
    assert NP.abs(wts - m["bhw_freq_wts"]).max() <= 1e-13
    out = {}
    for name, x in (("skyvis_lag", m["skyvis_freq"]), ("vis_lag", m["vis_freq"]), ("vis_noise_lag", m["vis_noise_freq"]),
                    ("lag_kernel", NP.ones_like(m["bp"]))):
        out[name], lags, corr = O.subband_delay_transform(x, m["bp"], wts, df, pad=1.0)
        ref = m["bhw_" + name]
        assert out[name].shape == ref.shape and NP.abs(out[name] - ref).max() <= 1e-12 * NP.abs(ref).max(), name
    assert NP.allclose(lags, m["bhw_lags"], rtol=0, atol=1e-18) and NP.allclose(corr, m["bhw_lag_corr_length"], rtol=1e-12)
    rl, rk, rs, rc = O.subband_resample(lags, out["lag_kernel"], [out["skyvis_lag"], out["vis_lag"], out["vis_noise_lag"]], m["bw_eff"],
                                        lags.size * df)
    assert NP.allclose(rl, m["bhw_rs_lags"], rtol=0, atol=1e-18) and NP.allclose(rc, m["bhw_rs_lag_corr_length"], rtol=1e-12)
    assert NP.allclose(rk, m["bhw_rs_lag_kernel"], rtol=1e-11, atol=1e-12 * NP.nanmax(NP.abs(rk)), equal_nan=True)
    for got, name in zip(rs, ("skyvis_lag", "vis_lag", "vis_noise_lag")):
        ref = m["bhw_rs_" + name]
        assert got.shape == ref.shape and NP.abs(got - ref).max() <= 1e-11 * NP.abs(ref).max(), name
    # rectangular window, no padding
    wr = O.subband_weights(m["chans"], [1.3e6], [m["chans"][15]], "rect")
    assert NP.abs(wr - m["rect_freq_wts"]).max() <= 1e-13
    lag, lags0, _ = O.subband_delay_transform(m["skyvis_freq"], m["bp"], wr, df, pad=0.0)
    assert NP.abs(lag - m["rect_skyvis_lag"]).max() <= 1e-12 * NP.abs(lag).max()


ROI_CASES = [("zenith_achromatic", {"radius": None, "center": None, "center_coords": None}),
             ("zenith_r30_chromatic", {"radius": 30.0, "center": None, "center_coords": None, "pbeam_chromaticity": True}),
             ("offzenith_altaz_reffreq", {"radius": 25.0, "center": NP.asarray([60.0, 140.0]), "center_coords": "altaz", "pbeam_reffreq": 151.3e6}),
             ("offzenith_hadec", {"radius": 40.0, "center": NP.asarray([20.0, -10.0]), "center_coords": "hadec"}),
             ("given_ind", {"ind": NP.asarray([3, 17, 44, 120, 250]), "radius": 90.0}),
             ("given_ind_pbeam", {"ind": NP.asarray([5, 6, 7])})]
ROI_TELESCOPE = {"id": "hera", "shape": "dish", "size": 14.0, "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz", "groundplane": None}


def test_roi_parameters_append_settings_matches_reference():
    """ROI_parameters.append_settings (interferometry.py:4221-4617) run by the reference: indices and beam tables."""
    g = _load("roi_parameters.npz")
    lat = float(g["latitude"])
    pinfo = {"pointing_center": NP.asarray([[90.0, 270.0]]), "pointing_coords": "altaz"}
    for name, ri in ROI_CASES:
        ri = dict(ri)
        if name == "given_ind_pbeam":
            ri["pbeam"] = g["pbeam_in_" + name]
        ind, pbeam, radius, center = O.roi_append_settings(g["hadec"], "hadec", lat, g["freq"], dict(ROI_TELESCOPE), ri, pinfo=pinfo, lst=10.0)
        assert NP.array_equal(ind, g["ind_" + name]), name
        assert pbeam.shape == g["pbeam_" + name].shape and pbeam.dtype == g["pbeam_" + name].dtype, name
        assert NP.allclose(pbeam, g["pbeam_" + name], rtol=1e-11, atol=1e-14), name
        if "radius_" + name in g.files and radius is not None:
            assert float(g["radius_" + name]) == radius
    assert int(g["n_entries"]) == len(ROI_CASES) + 1 and str(g["center_coords"]) == "altaz"
