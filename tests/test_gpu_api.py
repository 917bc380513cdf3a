"""GPU tests of the reference-facing Python surface (InterferometerArray / DelaySpectrum /
geometric_delay) against the golden vectors produced by the reference's own code and against the
oracle."""
import os

import numpy as NP
import pytest
import torch

from oracle import prisim_oracle as O
from tests.test_oracle_golden import OBSERVE_CASES

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
TOL = 1e-5


def rel_err(Vg, Vo):
    rms_b = NP.sqrt(NP.mean(NP.abs(Vo) ** 2, axis=tuple(range(1, Vo.ndim)), keepdims=True))
    return float((NP.abs(Vg - Vo) / NP.where(rms_b > 0, rms_b, 1.0)).max())


@pytest.mark.parametrize("tag", sorted(OBSERVE_CASES))
def test_observe_noise_delay_against_reference_golden(tag):
    """Replays the calls make_golden.py made on the reference's InterferometerArray."""
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    from prisim_b200.skymodel import SkyModel
    from prisim_b200.delay_spectrum import DelaySpectrum
    g = NP.load(os.path.join(GOLD, "observe_{0}.npz".format(tag)))
    case = OBSERVE_CASES[tag]
    nbl, nchan, nsnap = g["skyvis_freq"].shape
    labels = [("B{0}".format(i), "A{0}".format(i)) for i in range(nbl)]
    ia = InterferometerArray(labels, g["bl"], g["chans"], telescope=dict(case["telescope"]), eff_Q=0.96, latitude=float(g["latitude"]),
                             longitude=21.4278, altitude=0.0, skycoords="hadec", A_eff=154.0 * 0.65,
                             pointing_coords=case.get("pointing_coords", "hadec"), baseline_coords="localenu", freq_scale="Hz",
                             device=0, noise_seed=5)
    Tsysinfo = {"Trx": 50.0, "Tant": {"T0": 200.0, "f0": 150e6, "spindex": -2.55}, "Tnet": None}
    nsrc0 = g["flux"].size
    for j in range(nsnap):
        parms = {"location": g["hadec_{0}".format(j)], "coords": "hadec", "spec_type": "func", "frequency": [150e6],
                 "spec_parms": {"name": NP.repeat("power-law", nsrc0), "power-law-index": g["spindex"], "freq-ref": NP.full(nsrc0, 150e6),
                                "flux-scale": g["flux"]}}
        if g["src_shape"].size:
            parms["src_shape"] = g["src_shape"]
        kw = {}
        if "roi_ind_{0}".format(j) in g.files:
            kw["roi_info"] = {"ind": g["roi_ind_{0}".format(j)], "pbeam": g["roi_pbeam_{0}".format(j)]}
        ia.observe(SimpleTime(2451545.0 + j * 0.01, float(g["lsts"][j])), Tsysinfo, g["bandpass"], g["pointing"], SkyModel(init_parms=parms),
                   float(g["t_acc"][j]), roi_radius=case.get("roi_radius", None),
                   gradient_mode="baseline" if "gradient_baseline" in g.files else None, **kw)
        assert NP.array_equal(ia.obs_catalog_indices[j], g["m2_{0}".format(j)])
    assert ia.n_acc == int(g["n_acc"]) and NP.isclose(ia.t_obs, float(g["t_obs"]))
    assert NP.allclose(ia.pointing_center, g["pointing_center"])
    assert rel_err(ia.skyvis_freq, g["skyvis_freq"]) <= TOL
    assert NP.allclose(ia.Tsys, g["Tsys"]) and NP.allclose(ia.bp, g["bp"])
    if "gradient_baseline" in g.files:       # gradient_mode='baseline' (:6312-6343) and apply_gradients (:6726-6819)
        G = ia.gradient["baseline"]
        assert G.shape == g["gradient_baseline"].shape and ia.gradient_mode == "baseline"
        for i in range(3):
            assert rel_err(G[i], g["gradient_baseline"][i]) <= TOL
        dV = ia.apply_gradients(perturbations={"baseline": g["perturbations"].copy()})
        assert dV.shape == g["delta_skyvis_freq"].shape
        assert rel_err(dV.reshape((-1,) + dV.shape[2:]), g["delta_skyvis_freq"].reshape((-1,) + dV.shape[2:])) <= TOL
        with pytest.raises(KeyError):
            ia.apply_gradients(gradient_mode="baseline", perturbations={"skypos": NP.zeros((3, nbl))})
        with pytest.raises(ValueError):
            ia.apply_gradients(perturbations={"baseline": NP.zeros((3, nbl + 1))})
    else:
        assert ia.gradient == {}
        with pytest.raises(AttributeError):
            ia.apply_gradients(perturbations={"baseline": NP.zeros((3, nbl))})
    ia.generate_noise()
    ia.add_noise()
    assert NP.allclose(ia.vis_rms_freq, g["vis_rms_freq"], rtol=1e-12)
    z = ia.vis_noise_freq / (g["vis_rms_freq"] / NP.sqrt(2))
    assert abs(z.real.std() - 1) < 0.1 and abs(z.imag.std() - 1) < 0.1      # statistical parity (different RNG)
    assert NP.allclose(ia.vis_freq, ia.skyvis_freq + ia.vis_noise_freq)
    for pad, key in ((1.0, "skyvis_lag"), (0.0, "skyvis_lag_pad0"), (0.5, "skyvis_lag_pad05")):
        ia.delay_transform(pad=pad, freq_wts=g["window"], verbose=False)
        assert ia.skyvis_lag.shape == g[key].shape
        assert rel_err(ia.skyvis_lag, g[key]) <= TOL
        if pad == 1.0:
            assert NP.allclose(ia.lags, g["lags"])
            assert rel_err(ia.lag_kernel, g["lag_kernel"]) <= 1e-10
            lag_o, _ = O.delay_transform(ia.vis_freq, g["bp"], O.broadcast_freq_wts(g["window"], nbl, nchan, nsnap),
                                         g["chans"][1] - g["chans"][0], pad=1.0)
            assert rel_err(ia.vis_lag, lag_o) <= 1e-10
    ds = DelaySpectrum(interferometer_array=ia)
    res = ds.delay_transform(pad=1.0, freq_wts=g["window"], action="store", verbose=False)
    assert set(res) == {"freq_wts", "pad", "lags", "vis_lag", "skyvis_lag", "vis_noise_lag", "lag_kernel"}
    assert rel_err(res["skyvis_lag"], g["skyvis_lag"]) <= TOL and NP.allclose(res["lags"], g["lags"])
    res2 = ds.delay_transform(pad=1.0, freq_wts=g["window"], downsample=False, verbose=False)
    assert res2["skyvis_lag"].shape[1] == 2 * nchan and res2["lags"].size == 2 * nchan
    assert rel_err(res2["skyvis_lag"][:, ::2], g["skyvis_lag"]) <= TOL


def test_gradient_point_sources_vs_oracle_and_finite_difference():
    """gradient_mode='baseline' on a point-source sky (the reference's own branch is unreachable there, :6263/:6343):
    against the oracle, and against a finite difference of the GPU visibilities themselves."""
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    cfg = S.config1(nsnap=1)
    mk = lambda bl: InterferometerArray(cfg["labels"], bl, cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                                        skycoords="radec", pointing_coords="hadec", device=0)
    args = (SimpleTime(2451545.0, 0.0), {"Tnet": 300.0}, NP.ones(cfg["channels"].size), cfg["pointing_hadec"], cfg["skymodel"], cfg["t_acc"])
    ia = mk(cfg["baselines"])
    ia.precision = "fp64"
    ia.observe(*args, gradient_mode="baseline")
    G = ia.gradient["baseline"][..., 0]
    sm = cfg["skymodel"]
    hadec = NP.stack(((0.0 - sm.location[:, 0]) % 360.0, sm.location[:, 1]), axis=1)
    sp = sm.spec_parms
    (Vo, Go), _ = O.observe_snapshot(cfg["baselines"], cfg["channels"], hadec, "hadec", cfg["latitude"], cfg["pointing_hadec"], "hadec",
                                     dict(cfg["telescope"]), sp["flux-scale"], sp["power-law-index"], sp["freq-ref"], gradient=True)
    assert rel_err(ia.skyvis_freq[..., 0], Vo) <= 1e-9
    for i in range(3):
        assert rel_err(G[i], Go[i]) <= 1e-9
    # fp32 path within tolerance of the fp64 one
    ib = mk(cfg["baselines"])
    ib.observe(*args, gradient_mode="baseline")
    for i in range(3):
        assert rel_err(ib.gradient["baseline"][i, ..., 0], Go[i]) <= TOL
    # finite difference along east: V(b + h e_x) - V(b - h e_x) = 2 dV, the phase-centre (zenith) term does not move with b_x
    h = 1e-3
    dbl = NP.zeros_like(cfg["baselines"]); dbl[:, 0] = h
    Vp, Vm = mk(cfg["baselines"] + dbl), mk(cfg["baselines"] - dbl)
    for a in (Vp, Vm):
        a.precision = "fp64"
        a.observe(*args)
    pert = NP.zeros((1, 3, cfg["baselines"].shape[0])); pert[0, 0, :] = h
    dV = ia.apply_gradients(perturbations={"baseline": pert})[0, :, :, 0]
    fd = 0.5 * (Vp.skyvis_freq[..., 0] - Vm.skyvis_freq[..., 0])
    assert NP.abs(fd - dV).max() <= 1e-6 * NP.abs(dV).max()


def test_tsys_bandpass_bookkeeping_forms_against_reference_golden():
    """Every bandpass / Tsysinfo / bpcorrect form observe() accepts (interferometry.py:5993-6086), replaying the reference's run:
    Tsys, bp, thermal rms, visibilities and the bandpass-weighted delay spectra."""
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    from prisim_b200.skymodel import SkyModel
    g = NP.load(os.path.join(GOLD, "bookkeeping.npz"))
    nbl, nchan, nsnap = g["Tsys"].shape
    lat = float(g["latitude"])
    ia = InterferometerArray([("B{0}".format(i), "A{0}".format(i)) for i in range(nbl)], g["bl"], g["chans"], telescope=dict(OBSERVE_CASES["hera"]["telescope"]),
                             eff_Q=0.9, latitude=lat, skycoords="hadec", A_eff=g["A_eff"], pointing_coords="hadec", freq_scale="Hz", device=0)
    nsrc0 = g["flux"].size
    sky = SkyModel(init_parms={"location": g["hadec"], "coords": "hadec", "spec_type": "func", "frequency": [150e6],
                               "spec_parms": {"name": NP.repeat("power-law", nsrc0), "power-law-index": g["spindex"], "freq-ref": NP.full(nsrc0, 150e6),
                                              "flux-scale": g["flux"]}})
    tant = {"Trx": 40.0, "Tant": {"T0": 180.0, "f0": 150e6, "spindex": -2.5}}
    forms = [(g["bp1"], {"Tnet": 150.0}, None),
             (g["bp2"], {"Tnet": NP.linspace(100.0, 200.0, nbl)}, None),
             (g["bp3"], dict(tant, Tnet=None), g["bc2"]),
             (g["bp1"], {"Tnet": NP.linspace(90.0, 120.0, nchan)}, g["bc3"]),
             (g["bp1"], dict(tant), g["bc4"])]
    for j, (bp, ts, bc) in enumerate(forms):
        ia.observe(SimpleTime(2451545.0 + j * 0.01, 5.0 * j), ts, bp, NP.asarray([0.0, lat]), sky, float(g["t_acc"][j]), bpcorrect=bc)
    assert NP.allclose(ia.Tsys, g["Tsys"], rtol=1e-13) and NP.allclose(ia.bp, g["bp"], rtol=0, atol=0)
    assert NP.array_equal(ia.bp_wts, NP.ones((nbl, nchan, nsnap)))                 # unity until a delay transform sets them (:6024)
    assert rel_err(ia.skyvis_freq, g["skyvis_freq"]) <= TOL
    ia.generate_noise()
    assert NP.allclose(ia.vis_rms_freq, g["vis_rms_freq"], rtol=1e-12)
    ia.delay_transform(pad=1.0, freq_wts=g["window"], verbose=False)
    assert ia.bp_wts.shape == g["bp_wts"].shape and NP.allclose(ia.bp_wts, g["bp_wts"], rtol=1e-15)     # the window, broadcast (:8096-8106)
    assert rel_err(ia.skyvis_lag, g["skyvis_lag"]) <= TOL
    assert NP.abs(ia.lag_kernel - g["lag_kernel"]).max() <= 1e-10 * NP.abs(g["lag_kernel"]).max()
    with pytest.raises(ValueError):
        ia.observe(SimpleTime(2451545.1, 30.0), {"Tnet": 100.0}, g["bp1"], NP.asarray([0.0, lat]), sky, 10.0, bpcorrect=NP.ones(nchan + 1))
    with pytest.raises(KeyError):
        ia.observe(SimpleTime(2451545.1, 30.0), {"Trx": 40.0}, g["bp1"], NP.asarray([0.0, lat]), sky, 10.0)


def test_roi_about_the_pointing_centre_vs_oracle():
    """observe(roi_center='pointing_center', roi_radius=r) (interferometry.py:6212-6213): sources within r of an off-zenith
    pointing; the reference takes that branch through astropy coordinates, so the oracle gets the index list explicitly."""
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    from prisim_b200.skymodel import SkyModel
    g = NP.load(os.path.join(GOLD, "observe_gaussian_altazpointing.npz"))
    case = OBSERVE_CASES["gaussian_altazpointing"]
    nbl, nchan, _ = g["skyvis_freq"].shape
    lat = float(g["latitude"])
    ia = InterferometerArray([("B{0}".format(i), "A{0}".format(i)) for i in range(nbl)], g["bl"], g["chans"], telescope=dict(case["telescope"]),
                             latitude=lat, skycoords="hadec", pointing_coords="altaz", freq_scale="Hz", device=0)
    nsrc0 = g["flux"].size
    parms = {"location": g["hadec_0"], "coords": "hadec", "spec_type": "func", "frequency": [150e6],
             "spec_parms": {"name": NP.repeat("power-law", nsrc0), "power-law-index": g["spindex"], "freq-ref": NP.full(nsrc0, 150e6),
                            "flux-scale": g["flux"]}}
    radius = 35.0
    ia.observe(SimpleTime(2451545.0, float(g["lsts"][0])), {"Tnet": 100.0}, NP.ones(nchan), g["pointing"], SkyModel(init_parms=parms), 10.0,
               roi_radius=radius, roi_center="pointing_center")
    altaz = O.hadec2altaz(g["hadec_0"], lat, units="degrees")
    m2 = NP.where(O.sphdist(g["pointing"][1], g["pointing"][0], altaz[:, 1], altaz[:, 0]) <= radius)[0]
    assert 0 < m2.size < NP.sum(altaz[:, 0] >= 0) and NP.array_equal(ia.obs_catalog_indices[0], m2)
    pb = O.primary_beam_generator(altaz[m2], g["chans"] / 1e9, dict(case["telescope"]), skyunits="altaz", freq_scale="GHz", pointing_center=g["pointing"])
    Vo, _ = O.observe_snapshot(g["bl"], g["chans"], g["hadec_0"], "hadec", lat, g["pointing"], "altaz", dict(case["telescope"]), g["flux"], g["spindex"],
                               150e6, roi_info={"ind": m2, "pbeam": pb})
    assert rel_err(ia.skyvis_freq[..., 0], Vo) <= TOL


def test_delay_spectrum_allruns_against_reference_golden():
    """DelaySpectrum.delay_transform / delay_transform_allruns / horizon limits (delay_spectrum.py:1224-1342, :1475-1618,
    :2976-3030) replaying the reference's own DelaySpectrum on the 'hera' case."""
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    from prisim_b200.skymodel import SkyModel
    from prisim_b200.delay_spectrum import DelaySpectrum
    g = NP.load(os.path.join(GOLD, "observe_hera.npz"))
    a = NP.load(os.path.join(GOLD, "allruns_hera.npz"))
    nbl, nchan, nsnap = g["skyvis_freq"].shape
    ia = InterferometerArray([("B{0}".format(i), "A{0}".format(i)) for i in range(nbl)], g["bl"], g["chans"], telescope=dict(OBSERVE_CASES["hera"]["telescope"]),
                             eff_Q=0.96, latitude=float(g["latitude"]), longitude=21.4278, skycoords="hadec", A_eff=154.0 * 0.65,
                             pointing_coords="hadec", freq_scale="Hz", device=0, noise_seed=5)
    nsrc0 = g["flux"].size
    for j in range(nsnap):
        parms = {"location": g["hadec_{0}".format(j)], "coords": "hadec", "spec_type": "func", "frequency": [150e6],
                 "spec_parms": {"name": NP.repeat("power-law", nsrc0), "power-law-index": g["spindex"], "freq-ref": NP.full(nsrc0, 150e6),
                                "flux-scale": g["flux"]}}
        ia.observe(SimpleTime(2451545.0 + j * 0.01, float(g["lsts"][j])), {"Trx": 50.0, "Tant": {"T0": 200.0, "f0": 150e6, "spindex": -2.55}, "Tnet": None},
                   g["bandpass"], g["pointing"], SkyModel(init_parms=parms), float(g["t_acc"][j]))
    ia.generate_noise(); ia.add_noise()
    ia.delay_transform(pad=1.0, freq_wts=g["window"], verbose=False)     # as in the golden run: leaves the window in ia.bp_wts
    ds = DelaySpectrum(interferometer_array=ia)
    assert NP.allclose(ds.horizon_delay_limits, a["horizon_delay_limits"], rtol=1e-12, atol=1e-20)
    # external runs (the golden file's own arrays), vector window, pad 1 + downsample
    r = ds.delay_transform_allruns(a["runs"], pad=1.0, freq_wts=g["window"], downsample=True, verbose=False)
    assert r["vis_lag"].shape == a["vis_lag_pad1"].shape and r["lag_kernel"].shape == a["lag_kernel_pad1"].shape
    assert NP.abs(r["vis_lag"] - a["vis_lag_pad1"]).max() <= 1e-10 * NP.abs(a["vis_lag_pad1"]).max()
    assert NP.abs(r["lag_kernel"] - a["lag_kernel_pad1"]).max() <= 1e-10 * NP.abs(a["lag_kernel_pad1"]).max()
    assert NP.allclose(r["lags"], a["lags_pad1"]) and r["pad"] == 1.0 and r["freq_wts"].shape == (1, 1, 1, nchan, 1)
    # [nchan, nsnap] window, no padding
    r = ds.delay_transform_allruns(g["skyvis_freq"], pad=0.0, freq_wts=a["wts2"], verbose=False)
    assert NP.abs(r["vis_lag"] - a["vis_lag_pad0_w2"]).max() <= 1e-10 * NP.abs(a["vis_lag_pad0_w2"]).max()
    # freq_wts=None -> the weights stored by the last ia.delay_transform; pad 0.5 without downsampling (48-point transform)
    r = ds.delay_transform_allruns(g["skyvis_freq"], pad=0.5, freq_wts=None, downsample=False, verbose=False)
    assert r["vis_lag"].shape == a["vis_lag_pad05_full"].shape and NP.allclose(r["lags"], a["lags_pad05_full"])
    assert NP.abs(r["vis_lag"] - a["vis_lag_pad05_full"]).max() <= 1e-10 * NP.abs(a["vis_lag_pad05_full"]).max()
    assert NP.abs(r["lag_kernel"] - a["lag_kernel_pad05_full"]).max() <= 1e-10 * NP.abs(a["lag_kernel_pad05_full"]).max()
    # the object's own products
    r = ds.delay_transform(pad=1.0, freq_wts=g["window"], downsample=True, action="return", verbose=False)
    assert rel_err(r["skyvis_lag"], a["ds_skyvis_lag"]) <= TOL and NP.allclose(r["lags"], a["ds_lags"])
    assert NP.abs(r["lag_kernel"] - a["ds_lag_kernel"]).max() <= 1e-10 * NP.abs(a["ds_lag_kernel"]).max()
    with pytest.raises(ValueError):
        ds.delay_transform_allruns(g["skyvis_freq"][:, :-1, :])
    with pytest.raises(TypeError):
        ds.delay_transform_allruns([1, 2, 3])


def test_roi_parameters_against_reference_golden():
    """ROI_parameters.append_settings (interferometry.py:4221-4617) replaying the reference's own run; the tables then
    feed observe(roi_info=...) and give the same visibilities as the internal ROI + beam path."""
    from prisim_b200.interferometry import ROI_parameters, InterferometerArray, SimpleTime
    from prisim_b200.skymodel import SkyModel
    from tests.test_oracle_golden import ROI_CASES, ROI_TELESCOPE
    g = NP.load(os.path.join(GOLD, "roi_parameters.npz"))
    lat = float(g["latitude"])
    nsrc0 = g["hadec"].shape[0]
    sky = SkyModel(init_parms={"location": g["hadec"], "coords": "hadec", "spec_type": "func", "frequency": [150e6],
                               "spec_parms": {"name": NP.repeat("power-law", nsrc0), "power-law-index": NP.zeros(nsrc0),
                                              "freq-ref": NP.full(nsrc0, 150e6), "flux-scale": NP.ones(nsrc0)}})
    tel = dict(ROI_TELESCOPE); tel.update(latitude=lat, longitude=21.4278, altitude=0.0)
    roi = ROI_parameters(device=0)
    for name, ri in ROI_CASES:
        ri = dict(ri)
        if name == "given_ind_pbeam":
            ri["pbeam"] = g["pbeam_in_" + name]
        roi.append_settings(sky, g["freq"], pinfo={"pointing_center": NP.asarray([[90.0, 270.0]]), "pointing_coords": "altaz"}, lst=10.0,
                            time_jd=2451545.0, roi_info=ri, telescope=tel, freq_scale="Hz")
        assert NP.array_equal(roi.info["ind"][-1], g["ind_" + name]), name
        pb = roi.info["pbeam"][-1]
        assert pb.shape == g["pbeam_" + name].shape and pb.dtype == g["pbeam_" + name].dtype, name
        assert NP.abs(pb - g["pbeam_" + name]).max() <= 2e-7, name                     # device beam table is fp32
    roi.append_settings(None, g["freq"], telescope=tel, freq_scale="Hz")
    assert len(roi.info["ind"]) == int(g["n_entries"]) and roi.info["center_coords"] == str(g["center_coords"])
    assert NP.allclose(NP.concatenate([NP.asarray(c, dtype=float).reshape(1, 2) for c in roi.info["center"]]), g["centers"])
    with pytest.raises(ValueError):
        roi.append_settings(sky, g["freq"], pinfo=None, roi_info={"radius": 10.0, "center": None}, telescope=tel)
    with pytest.raises(ValueError):
        roi.append_settings(sky, g["freq"], roi_info={"ind": NP.arange(4), "pbeam": NP.ones((3, g["freq"].size))}, telescope=tel)
    # the tables drive observe(): same visibilities as the internal cull + beam (zenith ROI, chromatic table)
    bl = NP.asarray([[14.6, 0.0, 0.0], [0.0, 29.2, 0.0], [43.8, 14.6, 0.0]])
    mk = lambda: InterferometerArray(["a", "b", "c"], bl, g["freq"], telescope=dict(ROI_TELESCOPE), latitude=lat, skycoords="hadec",
                                     pointing_coords="altaz", freq_scale="Hz", device=0)
    args = (SimpleTime(2451545.0, 10.0), {"Tnet": 100.0}, NP.ones(g["freq"].size), NP.asarray([90.0, 270.0]), sky, 10.0)
    ia, ib = mk(), mk()
    ia.observe(*args, roi_radius=30.0)
    ib.observe(*args, roi_info={"ind": roi.info["ind"][1], "pbeam": roi.info["pbeam"][1]})
    assert NP.array_equal(ia.obs_catalog_indices[0], ib.obs_catalog_indices[0])
    assert rel_err(ib.skyvis_freq, ia.skyvis_freq) <= 1e-6


def test_multi_window_delay_transform_against_reference_golden():
    """Sub-band delay transforms (interferometry.py:8141-8287) replaying the reference's own run on the 'hera' case."""
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    from prisim_b200.skymodel import SkyModel
    g = NP.load(os.path.join(GOLD, "observe_hera.npz"))
    m = NP.load(os.path.join(GOLD, "multiwin_hera.npz"))
    nbl, nchan, nsnap = g["skyvis_freq"].shape
    ia = InterferometerArray([("B{0}".format(i), "A{0}".format(i)) for i in range(nbl)], g["bl"], g["chans"], telescope=dict(OBSERVE_CASES["hera"]["telescope"]),
                             eff_Q=0.96, latitude=float(g["latitude"]), longitude=21.4278, skycoords="hadec", A_eff=154.0 * 0.65,
                             pointing_coords="hadec", freq_scale="Hz", device=0, noise_seed=5)
    nsrc0 = g["flux"].size
    for j in range(nsnap):
        parms = {"location": g["hadec_{0}".format(j)], "coords": "hadec", "spec_type": "func", "frequency": [150e6],
                 "spec_parms": {"name": NP.repeat("power-law", nsrc0), "power-law-index": g["spindex"], "freq-ref": NP.full(nsrc0, 150e6),
                                "flux-scale": g["flux"]}}
        ia.observe(SimpleTime(2451545.0 + j * 0.01, float(g["lsts"][j])), {"Trx": 50.0, "Tant": {"T0": 200.0, "f0": 150e6, "spindex": -2.55}, "Tnet": None},
                   g["bandpass"], g["pointing"], SkyModel(init_parms=parms), float(g["t_acc"][j]))
    res0 = ia.multi_window_delay_transform(m["bw_eff"], freq_center=m["freq_center"], shape="bhw", verbose=False)   # before noise exists
    assert set(res0) == {"skyvis_lag", "lag_kernel", "lag_corr_length"}
    ia.generate_noise()
    df = g["chans"][1] - g["chans"][0]
    for shape, bw, fc, sfx, pad in (("bhw", m["bw_eff"], m["freq_center"], "pad1", 1.0), ("bhw", m["bw_eff"], m["freq_center"], "pad0", 0.0),
                                    ("rect", 1.2e6, None, "rect", 1.0)):
        res = ia.multi_window_delay_transform(bw, freq_center=fc, shape=shape, pad=pad, verbose=False)
        assert res["skyvis_lag"].shape == m["skyvis_lag_" + sfx].shape
        nw = res["skyvis_lag"].shape[1]
        for i in range(nw):
            assert rel_err(res["skyvis_lag"][:, i], m["skyvis_lag_" + sfx][:, i]) <= TOL
        assert NP.allclose(res["lag_corr_length"], m["lag_corr_length_" + sfx], rtol=1e-12)
        assert NP.array_equal(ia.subband_windows(bw, fc, shape), O.multi_window_weights(g["chans"], bw, fc, shape))
        if sfx != "rect":
            assert NP.abs(res["lag_kernel"] - m["lag_kernel_" + sfx]).max() <= 1e-10 * NP.abs(m["lag_kernel_" + sfx]).max()
            nz_o, _ = O.multi_window_delay_transform(ia.vis_noise_freq, g["bp"], O.multi_window_weights(g["chans"], bw, fc, shape), df, pad=pad)
            assert NP.abs(res["vis_noise_lag"] - nz_o).max() <= 1e-10 * NP.abs(nz_o).max()
    with pytest.raises(ValueError):
        ia.multi_window_delay_transform(-1.0)
    with pytest.raises(ValueError):
        ia.multi_window_delay_transform(1e6, freq_center=g["chans"][-1])
    with pytest.raises(ValueError):
        ia.multi_window_delay_transform(1e6, shape="hann")
    with pytest.raises(TypeError):
        ia.multi_window_delay_transform("wide")


def test_subband_delay_transform_against_reference_golden():
    """DelaySpectrum.subband_delay_transform (delay_spectrum.py:1842-2248) against the reference's own run: full-resolution and
    resampled products, window bookkeeping, argument checks."""
    from prisim_b200.delay_spectrum import DelaySpectrum
    from prisim_b200.interferometry import InterferometerArray
    g = NP.load(os.path.join(GOLD, "observe_hera.npz"))
    m = NP.load(os.path.join(GOLD, "subband_hera.npz"))
    chans = m["chans"]
    nbl, nchan, nsnap = m["skyvis_freq"].shape
    from prisim_b200 import engine
    ia = InterferometerArray([("B{0}".format(i), "A{0}".format(i)) for i in range(nbl)], g["bl"], chans, telescope=dict(OBSERVE_CASES["hera"]["telescope"]),
                             latitude=float(g["latitude"]), skycoords="hadec", pointing_coords="hadec", freq_scale="Hz", device=0, noise_seed=5)
    # the reference's own frequency-domain products, so that the comparison isolates the sub-band transform
    for t in range(nsnap):
        ia._skyvis.append(engine._c128(NP.ascontiguousarray(m["skyvis_freq"][:, :, t]), 0)); ia._vis.append(engine._c128(NP.ascontiguousarray(m["vis_freq"][:, :, t]), 0))
        ia._noise.append(engine._c128(NP.ascontiguousarray(m["vis_noise_freq"][:, :, t]), 0)); ia._bp.append(engine._f64(NP.ascontiguousarray(m["bp"][:, :, t]), 0))
    ds = DelaySpectrum(interferometer_array=ia)
    args = dict(freq_center={"sim": m["freq_center"]}, shape={"sim": "bhw"}, pad={"sim": 1.0}, verbose=False)
    r = ds.subband_delay_transform({"sim": m["bw_eff"]}, action="return_oversampled", **args)["sim"]
    assert NP.abs(r["freq_wts"] - m["bhw_freq_wts"]).max() <= 1e-13 and NP.allclose(r["lags"], m["bhw_lags"], rtol=0, atol=1e-18)
    assert NP.allclose(r["lag_corr_length"], m["bhw_lag_corr_length"], rtol=1e-12) and r["npad"] == nchan
    for name in ("skyvis_lag", "vis_lag", "vis_noise_lag", "lag_kernel"):
        ref = m["bhw_" + name]
        assert r[name].shape == ref.shape and NP.abs(r[name] - ref).max() <= 1e-11 * NP.abs(ref).max(), name
    rs = ds.subband_delay_spectra_resampled["sim"]
    assert NP.allclose(rs["lags"], m["bhw_rs_lags"], rtol=0, atol=1e-18) and NP.allclose(rs["lag_corr_length"], m["bhw_rs_lag_corr_length"], rtol=1e-12)
    assert NP.allclose(rs["lag_kernel"], m["bhw_rs_lag_kernel"], rtol=1e-10, atol=1e-11 * NP.nanmax(NP.abs(m["bhw_rs_lag_kernel"])), equal_nan=True)
    for name in ("skyvis_lag", "vis_lag", "vis_noise_lag"):
        ref = m["bhw_rs_" + name]
        assert rs[name].shape == ref.shape and NP.abs(rs[name] - ref).max() <= 1e-10 * NP.abs(ref).max(), name
    r2 = ds.subband_delay_transform({"sim": NP.asarray([1.3e6])}, freq_center={"sim": NP.asarray([chans[15]])}, pad={"sim": 0.0},
                                    action="return_resampled", verbose=False)["sim"]
    assert NP.abs(ds.subband_delay_spectra["sim"]["freq_wts"] - m["rect_freq_wts"]).max() <= 1e-13
    assert NP.abs(ds.subband_delay_spectra["sim"]["skyvis_lag"] - m["rect_skyvis_lag"]).max() <= 1e-11 * NP.abs(m["rect_skyvis_lag"]).max()
    # (this window sits mid-band, so its Fourier-resampled spectrum is numerically zero in the reference as well: absolute tolerance)
    assert NP.abs(r2["skyvis_lag"] - m["rect_rs_skyvis_lag"]).max() <= 1e-12 * NP.abs(m["rect_skyvis_lag"]).max()
    assert NP.allclose(r2["lags"], m["rect_rs_lags"], rtol=0, atol=1e-18)
    with pytest.raises(TypeError):
        ds.subband_delay_transform(1e6)
    with pytest.raises(ValueError):
        ds.subband_delay_transform({"sim": [1e6, 2e6]}, freq_center={"sim": [chans[5], chans[6], chans[7]]})
    with pytest.raises(ValueError):
        ds.subband_delay_transform({"sim": 1e6}, freq_center={"sim": chans[0]})
    with pytest.raises(NotImplementedError):
        ds.subband_delay_transform({"sim": 1e6}, fftpow={"sim": 2.0})


def test_duplicate_measurements_against_reference_golden():
    """Unique baselines -> redundant sets (interferometry.py:6823-6907), replaying the reference's own run."""
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    from prisim_b200.skymodel import SkyModel
    from tests.test_oracle_golden import _dup_groups
    g = NP.load(os.path.join(GOLD, "duplicate.npz"))
    dt = [("A2", "U1"), ("A1", "U1")]
    labels = NP.asarray([tuple(l) for l in g["ulabels"].tolist()], dtype=dt)
    groups = {k: NP.asarray(v, dtype=dt) for k, v in _dup_groups(g).items()}
    hera = {"id": "hera", "shape": "dish", "size": 14.0, "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz", "groundplane": None}
    lat = float(g["latitude"])
    ia = InterferometerArray(labels, g["bl"], g["chans"], telescope=hera, eff_Q=0.96, latitude=lat, longitude=21.4278, skycoords="hadec",
                             A_eff=154.0 * 0.65, pointing_coords="hadec", freq_scale="Hz", blgroupinfo={"groups": groups, "reversemap": None},
                             device=0, noise_seed=3)
    nsrc0, nchan = g["flux"].size, g["chans"].size
    for j in range(2):
        parms = {"location": g["hadec_{0}".format(j)], "coords": "hadec", "spec_type": "func", "frequency": [150e6],
                 "spec_parms": {"name": NP.repeat("power-law", nsrc0), "power-law-index": g["spindex"], "freq-ref": NP.full(nsrc0, 150e6),
                                "flux-scale": g["flux"]}}
        ia.observe(SimpleTime(2451545.0 + j * 0.01, float(g["lsts"][j])), {"Trx": 50.0, "Tant": {"T0": 200.0, "f0": 150e6, "spindex": -2.55}, "Tnet": None},
                   NP.ones(nchan), NP.asarray([0.0, lat]), SkyModel(init_parms=parms), 10.7, gradient_mode="baseline")
    assert rel_err(ia.skyvis_freq, g["skyvis_unique"]) <= TOL
    ia.project_baselines(ref_point={"location": NP.asarray([[0.0, lat]]), "coords": "hadec"})
    assert NP.allclose(ia.projected_baselines, g["projected_unique"], atol=1e-9)
    unique = ia.skyvis_freq.copy()
    grad_unique = ia.gradient["baseline"].copy()
    ia.duplicate_measurements()
    assert [tuple(l) for l in ia.labels.tolist()] == [tuple(l) for l in g["labels_out"].tolist()]
    assert NP.array_equal(ia.baselines, g["baselines_out"]) and NP.allclose(ia.baseline_lengths, g["baseline_lengths_out"])
    assert NP.allclose(ia.projected_baselines, g["projected_out"], atol=1e-9)
    num = [3, 2, 1, 1]
    assert NP.array_equal(ia.skyvis_freq, NP.repeat(unique, num, axis=0)) and rel_err(ia.skyvis_freq, g["skyvis_out"]) <= TOL
    assert NP.array_equal(ia.gradient["baseline"], NP.repeat(grad_unique, num, axis=1))
    assert NP.allclose(ia.Tsys, g["Tsys_out"]) and NP.allclose(ia.bp, g["bp_out"])
    assert NP.allclose(ia.vis_rms_freq, g["vis_rms_out"], rtol=1e-12)
    assert ia.vis_noise_freq.shape == tuple(g["vis_noise_shape"]) and NP.allclose(ia.vis_freq - ia.vis_noise_freq, ia.skyvis_freq)
    # redundant copies share the sky signal but not the noise
    assert NP.abs(ia.vis_noise_freq[0] - ia.vis_noise_freq[1]).min() > 0
    # the delay transform runs on the expanded set
    ia.delay_transform(pad=1.0, verbose=False)
    assert ia.skyvis_lag.shape == (7, nchan, 2)
    ia.duplicate_measurements()                                                        # complete now: no-op (:6852-6857)
    assert ia.baselines.shape[0] == 7


def test_config1_observing_run_vs_oracle():
    """BASELINE config 1 in full through observing_run (drift scan, 10 snapshots)."""
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray
    cfg = S.config1()
    ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", pointing_coords="hadec", device=0)
    nsnap, t_acc = cfg["nsnap"], cfg["t_acc"]
    ia.observing_run(cfg["pointing_hadec"], cfg["skymodel"], t_acc, nsnap * t_acc, cfg["channels"], NP.ones(128), 300.0, 0.0,
                     mode="drift", pointing_coords="hadec", verbose=False)
    assert ia.n_acc == nsnap and ia.skyvis_freq.shape == (171, 128, nsnap)
    import tempfile
    with tempfile.TemporaryDirectory() as td:                      # NPZ product, same keys as interferometry.py:8859-8863
        ia.save(os.path.join(td, "simvis"), fmt="NPZ", verbose=False)
        z = NP.load(os.path.join(td, "simvis.npz"))
        assert set(z.files) == {"skyvis_freq", "lst", "freq", "timestamp", "bl", "bl_length"}
        assert NP.array_equal(z["skyvis_freq"], ia.skyvis_freq) and NP.allclose(z["freq"], cfg["channels"])
        with pytest.raises(NotImplementedError):
            ia.save(os.path.join(td, "simvis"), fmt="HDF5")
    sky = cfg["skymodel"]; sp = sky.spec_parms
    Vo = []
    for j in range(nsnap):
        lst = (0.0 + (t_acc / 3.6e3) * j) * 15.0
        assert NP.isclose(ia.lst[j], lst)
        hadec = NP.stack((lst - sky.location[:, 0], sky.location[:, 1]), 1)
        V, m2 = O.observe_snapshot(cfg["baselines"], cfg["channels"], hadec, "hadec", cfg["latitude"], cfg["pointing_hadec"], "hadec",
                                   cfg["telescope"], sp["flux-scale"], sp["power-law-index"], sp["freq-ref"])
        assert NP.array_equal(ia.obs_catalog_indices[j], m2)
        Vo.append(V)
    Vo = NP.stack(Vo, axis=2)
    assert rel_err(ia.skyvis_freq, Vo) <= TOL
    # redundant baselines see identical visibilities
    from prisim_b200.interferometry import uniq_baselines
    ub, first, counts, occ = uniq_baselines(cfg["baselines"])
    assert ub.shape[0] == 30
    grp = max(occ, key=len)                                        # b and -b share a group (conjugate visibilities): compare moduli
    assert len(grp) > 1
    assert NP.abs(NP.abs(ia.skyvis_freq[grp[1:]]) - NP.abs(ia.skyvis_freq[grp[0]])).max() <= 2e-5 * NP.abs(ia.skyvis_freq[grp[0]]).max()
    # track mode flips the bookkeeping to RA-Dec (interferometry.py:6620-6621)
    ib = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", pointing_coords="hadec", device=0)
    ib.observing_run(NP.asarray([10.0, -20.0]), sky, t_acc, 2 * t_acc, cfg["channels"], NP.ones(128), 300.0, 1.0, mode="track",
                     pointing_coords="radec", verbose=False)
    assert ib.pointing_coords == "radec" and ib.n_acc == 2
    lst1 = (1.0 + t_acc / 3.6e3) * 15.0
    hadec = NP.stack((lst1 - sky.location[:, 0], sky.location[:, 1]), 1)
    V, _ = O.observe_snapshot(cfg["baselines"], cfg["channels"], hadec, "hadec", cfg["latitude"], NP.asarray([10.0, -20.0]), "radec",
                              cfg["telescope"], sp["flux-scale"], sp["power-law-index"], sp["freq-ref"], lst=lst1)
    assert rel_err(ib.skyvis_freq[:, :, 1], V) <= TOL


def test_observe_input_errors_match_reference_types():
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    cfg = S.config1(nsrc=50, nchan=16, nsnap=1)
    ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", device=0)
    t = SimpleTime(2451545.0, 0.0)
    good = dict(Tsysinfo={"Tnet": 100.0}, bandpass=NP.ones(16), pointing_center=cfg["pointing_hadec"], skymodel=cfg["skymodel"], t_acc=10.0)
    with pytest.raises(ValueError):
        ia.observe(t, **dict(good, bandpass=NP.ones(15)))
    with pytest.raises(TypeError):
        ia.observe(t, **dict(good, Tsysinfo=300.0))
    with pytest.raises(KeyError):
        ia.observe(t, **dict(good, Tsysinfo={"Trx": 1.0}))
    with pytest.raises(ValueError):
        ia.observe(t, **dict(good, Tsysinfo={"Tnet": -5.0}))
    with pytest.raises(KeyError):
        ia.observe(t, roi_info={"ind": None}, **good)
    with pytest.raises(ValueError):
        ia.observe(t, roi_center="nadir", **good)
    with pytest.raises(TypeError):
        ia.observe(t, **dict(good, skymodel=object()))
    with pytest.raises(TypeError):
        ia.add_noise()
    assert ia.n_acc == 0
    with pytest.warns(UserWarning):                                         # empty ROI: zeros + warning (:6379)
        ia.observe(t, roi_radius=0.0, **good)
    assert ia.n_acc == 1 and NP.all(ia.skyvis_freq == 0)


def test_geometric_delay_and_mwa_config4_slice():
    from prisim_b200 import baseline_delay_horizon as DLY
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    g = NP.load(os.path.join(GOLD, "delays.npz"))
    lat = float(g["latitude"])
    assert NP.abs(DLY.geometric_delay(g["bl"], g["altaz"], altaz=True, hadec=False) - g["tau_altaz"]).max() < 1e-20
    assert NP.abs(DLY.geometric_delay(g["bl"], g["hadec"], hadec=True, latitude=lat) - g["tau_hadec"]).max() < 1e-20
    assert NP.allclose(DLY.horizon_delay_limits(g["bl"], O.altaz2dircos([90.0, 270.0])), g["horizon"], rtol=1e-12, atol=1e-22)
    with pytest.raises(ValueError):
        DLY.geometric_delay(g["bl"], g["hadec"], hadec=True)
    # BASELINE config 4 (MWA tile beam with quantised delays) on a slice the oracle can do
    cfg = S.config4(ntiles=24, nsrc=3000, nchan=96)
    ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", pointing_coords="altaz", device=0)
    ia.observe(SimpleTime(2451545.0, 50.0), {"Tnet": 200.0}, NP.ones(96), cfg["pointing_altaz"], cfg["skymodel"], cfg["t_acc"],
               pb_info=cfg["pb_info"])
    sky = cfg["skymodel"]; sp = sky.spec_parms
    hadec = NP.stack((50.0 - sky.location[:, 0], sky.location[:, 1]), 1)
    Vo, m2 = O.observe_snapshot(cfg["baselines"], cfg["channels"], hadec, "hadec", cfg["latitude"], cfg["pointing_altaz"], "altaz",
                                cfg["telescope"], sp["flux-scale"], sp["power-law-index"], sp["freq-ref"], pb_info=cfg["pb_info"])
    assert NP.array_equal(ia.obs_catalog_indices[0], m2)
    assert rel_err(ia.skyvis_freq[:, :, 0], Vo) <= TOL


def test_config3_diffuse_taper_slice_vs_oracle():
    """BASELINE config 3 shape (HEALPix diffuse sky, extended-source taper on) at nside 16.  The smooth sky
    cancels to ~1e-3 of its incoherent norm on the longer baselines; precision='auto' recomputes those in fp64."""
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    cfg = S.config3(nside=16, nchan=64, n_side=4, nsnap=2)
    ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", pointing_coords="hadec", device=0)
    sky = cfg["skymodel"]; sp = sky.spec_parms
    for j, lst in enumerate((0.0, 0.45)):
        ia.observe(SimpleTime(2451545.0 + j, lst), {"Tnet": 200.0}, NP.ones(64), cfg["pointing_hadec"], sky, cfg["t_acc"])
        hadec = NP.stack((lst - sky.location[:, 0], sky.location[:, 1]), 1)
        Vo, m2 = O.observe_snapshot(cfg["baselines"], cfg["channels"], hadec, "hadec", cfg["latitude"], cfg["pointing_hadec"], "hadec",
                                    cfg["telescope"], sp["flux-scale"], sp["power-law-index"], sp["freq-ref"], src_shape=sky.src_shape)
        assert NP.array_equal(ia.obs_catalog_indices[j], m2)
        assert rel_err(ia.skyvis_freq[:, :, j], Vo) <= TOL
    assert ia.precision_report[0]["fp64_baselines"] > 0
    # fp32-only is not enough here, fp64-only is
    for prec, ok in (("fp32", False), ("fp64", True)):
        ib = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                                 skycoords="radec", pointing_coords="hadec", device=0)
        ib.precision = prec
        ib.observe(SimpleTime(2451545.0, 0.0), {"Tnet": 200.0}, NP.ones(64), cfg["pointing_hadec"], sky, cfg["t_acc"])
        hadec = NP.stack((0.0 - sky.location[:, 0], sky.location[:, 1]), 1)
        Vo, _ = O.observe_snapshot(cfg["baselines"], cfg["channels"], hadec, "hadec", cfg["latitude"], cfg["pointing_hadec"], "hadec",
                                   cfg["telescope"], sp["flux-scale"], sp["power-law-index"], sp["freq-ref"], src_shape=sky.src_shape)
        assert (rel_err(ib.skyvis_freq[:, :, 0], Vo) <= TOL) == ok


def test_fp64_kernel_and_precision_control_on_point_sources():
    """fp64 kernel vs oracle at ~1e-12; 'auto' leaves a random point-source sky on the fp32 path except for the
    few baselines whose spectrum happens to cancel."""
    from prisim_b200 import engine, synthetic as S
    rng = NP.random.default_rng(8)
    nsrc, nchan = 1200, 160
    bl = S.array_baselines(S.hera_layout(4))[0] * 3.0
    freqs = 150e6 + (NP.arange(nchan) - nchan // 2) * 97656.25
    altaz = NP.stack((NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))), rng.uniform(0, 360, nsrc)), 1)
    dense = rng.uniform(0.05, 5.0, (nsrc, nchan))
    dircos, _ = engine.sky_cull(altaz, "altaz")
    Vo = O.skyvis_snapshot(bl, altaz, dense, freqs, NP.asarray([90.0, 270.0]))
    amp64 = engine.dense_to_amp_table(torch.as_tensor(dense).cuda(), dtype=torch.float64)
    V64 = engine.skyvis(dircos, amp64, nsrc, bl, (0, 0, 1.0), freqs, method="fp64").cpu().numpy()
    assert rel_err(V64, Vo) <= 1e-11
    fwhm = rng.uniform(0.05, 0.6, nsrc)
    Vt = engine.skyvis(dircos, amp64, nsrc, bl * 10, (0, 0, 1.0), freqs, src_fwhm_deg=engine._f64(fwhm, 0), method="fp64").cpu().numpy()
    Vto = O.skyvis_snapshot(bl * 10, altaz, dense, freqs, NP.asarray([90.0, 270.0]), src_shape=NP.stack((fwhm, fwhm, 0 * fwhm), 1))
    assert rel_err(Vt, Vto) <= 1e-6                                       # taper exponent coefficient is fp32
    amp32 = engine.dense_to_amp_table(torch.as_tensor(dense).cuda())
    V64b = engine.skyvis(dircos, amp32, nsrc, bl, (0, 0, 1.0), freqs, method="fp64").cpu().numpy()
    assert rel_err(V64b, Vo) <= 1e-6                                      # fp32 table: amplitude rounding only
    from prisim_b200._lib import PB200Error
    with pytest.raises(PB200Error):
        engine.skyvis(dircos, amp64, nsrc, bl, (0, 0, 1.0), freqs, method="recurrence")


def test_rotate_visibilities_against_reference_golden():
    """phase_centering / project_baselines / rotate_visibilities replayed on the GPU shim
    (interferometry.py:7655-7995; scripts/run_prisim.py:2282)."""
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    from prisim_b200.skymodel import SkyModel
    g = NP.load(os.path.join(GOLD, "observe_hera.npz"))
    r = NP.load(os.path.join(GOLD, "rotate_hera.npz"))
    case = OBSERVE_CASES["hera"]
    nbl, nchan, nsnap = g["skyvis_freq"].shape
    labels = [("B{0}".format(i), "A{0}".format(i)) for i in range(nbl)]
    ia = InterferometerArray(labels, g["bl"], g["chans"], telescope=dict(case["telescope"]), eff_Q=0.96, latitude=float(g["latitude"]),
                             skycoords="hadec", A_eff=154.0 * 0.65, pointing_coords="hadec", device=0)
    nsrc0 = g["flux"].size
    for j in range(nsnap):
        parms = {"location": g["hadec_{0}".format(j)], "coords": "hadec", "spec_type": "func", "frequency": [150e6],
                 "spec_parms": {"name": NP.repeat("power-law", nsrc0), "power-law-index": g["spindex"], "freq-ref": NP.full(nsrc0, 150e6),
                                "flux-scale": g["flux"]}}
        ia.observe(SimpleTime(2451545.0 + j * 0.01, float(g["lsts"][j])), {"Trx": 50.0, "Tant": {"T0": 200.0, "f0": 150e6, "spindex": -2.55}, "Tnet": None},
                   g["bandpass"], g["pointing"], SkyModel(init_parms=parms), float(g["t_acc"][j]))
    ia.generate_noise(); ia.add_noise()
    noise0, vis0 = ia.vis_noise_freq, ia.vis_freq
    ia.rotate_visibilities({"location": r["ref1"], "coords": "hadec"}, verbose=False)
    assert rel_err(ia.skyvis_freq, r["skyvis_rot1"]) <= TOL
    assert NP.allclose(ia.phase_center, r["pc_rot1"]) and ia.phase_center_coords == "hadec"
    assert NP.allclose(ia.projected_baselines, r["proj_rot1"], atol=1e-9)
    # the noise and noisy visibilities are rotated by the same phasor (:7876-7881)
    ph = r["skyvis_rot1"] / NP.where(g["skyvis_freq"] == 0, 1, g["skyvis_freq"])
    assert NP.abs(ia.vis_noise_freq - noise0 * ph).max() <= 1e-9 * NP.abs(noise0).max()
    assert NP.abs(ia.vis_freq - vis0 * ph).max() <= 1e-9 * NP.abs(vis0).max()
    ia.rotate_visibilities({"location": r["ref2"], "coords": "radec"}, verbose=False)
    assert rel_err(ia.skyvis_freq, r["skyvis_rot2"]) <= TOL
    assert NP.allclose(ia.phase_center, r["pc_rot2"]) and NP.allclose(ia.projected_baselines, r["proj_rot2"], atol=1e-9)
    with pytest.raises(KeyError):
        ia.rotate_visibilities({"location": r["ref1"]})
    with pytest.raises(TypeError):
        ia.phase_centering({"location": [1.0, 2.0], "coords": "hadec"})
    with pytest.raises(ValueError):
        ia.phase_centering({"location": NP.zeros((2, 2)), "coords": "hadec"})


def _subset_oracle(cfg, lst, bsel, csel, **kw):
    sky = cfg["skymodel"]; sp = sky.spec_parms
    hadec = NP.stack((lst - sky.location[:, 0], sky.location[:, 1]), 1)
    return O.observe_snapshot(cfg["baselines"][bsel], cfg["channels"][csel], hadec, "hadec", cfg["latitude"], kw.pop("pointing"),
                              kw.pop("pointing_coords"), cfg["telescope"], sp["flux-scale"], sp["power-law-index"], sp["freq-ref"], **kw)


def test_config3_full_size_snapshot_subset_parity():
    """BASELINE config 3 at full size for one snapshot: nside-256 diffuse sky (786,432 pixels, ~393k above the
    horizon, taper on) x HERA-331 (54,615 baselines) x 256 channels; parity on a baseline/channel subset."""
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    cfg = S.config3(nsnap=1)
    ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", pointing_coords="hadec", device=0)
    ia.observe(SimpleTime(2451545.0, 20.0), {"Tnet": 200.0}, NP.ones(256), cfg["pointing_hadec"], cfg["skymodel"], cfg["t_acc"])
    assert ia.baselines.shape[0] == 54615 and 380000 < ia.obs_catalog_indices[0].size < 400000
    rng = NP.random.default_rng(3)
    bsel = NP.sort(NP.concatenate(([0, 54614], rng.choice(54615, 6, replace=False))))
    csel = NP.sort(rng.choice(256, 24, replace=False))
    Vo, m2 = _subset_oracle(cfg, 20.0, bsel, csel, pointing=cfg["pointing_hadec"], pointing_coords="hadec", src_shape=cfg["skymodel"].src_shape)
    assert NP.array_equal(ia.obs_catalog_indices[0], m2)
    V = ia.skyvis_freq_device(0)
    Vg = V[torch.as_tensor(bsel).cuda()][:, torch.as_tensor(csel).cuda()].cpu().numpy()
    rms_b = V[torch.as_tensor(bsel).cuda()].abs().pow(2).mean(dim=1, keepdim=True).sqrt().cpu().numpy()
    assert float((NP.abs(Vg - Vo) / rms_b).max()) <= TOL
    assert ia.precision_report[0]["fp64_baselines"] > 0.5 * 54615          # the smooth sky cancels on most baselines


def test_config4_full_size_mwa_tile_subset_parity():
    """BASELINE config 4 at full size: 128 tiles (8,128 baselines) x 768 channels x 50k-source catalogue with the
    phased 4x4 tile beam (quantised delays) and ground plane."""
    from prisim_b200 import synthetic as S
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    cfg = S.config4()
    ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", pointing_coords="altaz", device=0)
    ia.observe(SimpleTime(2451545.0, 50.0), {"Tnet": 200.0}, NP.ones(768), cfg["pointing_altaz"], cfg["skymodel"], cfg["t_acc"], pb_info=cfg["pb_info"])
    assert ia.baselines.shape[0] == 8128
    rng = NP.random.default_rng(4)
    bsel = NP.sort(NP.concatenate(([0, 8127], rng.choice(8128, 10, replace=False))))
    csel = NP.sort(rng.choice(768, 32, replace=False))
    Vo, m2 = _subset_oracle(cfg, 50.0, bsel, csel, pointing=cfg["pointing_altaz"], pointing_coords="altaz", pb_info=cfg["pb_info"])
    assert NP.array_equal(ia.obs_catalog_indices[0], m2)
    V = ia.skyvis_freq_device(0)
    Vg = V[torch.as_tensor(bsel).cuda()][:, torch.as_tensor(csel).cuda()].cpu().numpy()
    rms_b = V[torch.as_tensor(bsel).cuda()].abs().pow(2).mean(dim=1, keepdim=True).sqrt().cpu().numpy()
    assert float((NP.abs(Vg - Vo) / rms_b).max()) <= TOL
    # every one of the 8,128 x 768 cells: the raw fp32 kernel against the fp64 kernel
    out = {}
    for prec in ("fp32", "fp64"):
        ib = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                                 skycoords="radec", pointing_coords="altaz", device=0)
        ib.precision = prec
        ib.observe(SimpleTime(2451545.0, 50.0), {"Tnet": 200.0}, NP.ones(768), cfg["pointing_altaz"], cfg["skymodel"], cfg["t_acc"], pb_info=cfg["pb_info"])
        out[prec] = ib.skyvis_freq_device(0)
    rms_all = out["fp64"].abs().pow(2).mean(dim=1, keepdim=True).sqrt()
    worst = float(((out["fp32"] - out["fp64"]).abs() / rms_all).max().item())
    print("config 4, all 8128 x 768 cells: max |dV|/rms_b = {0:.3e}".format(worst))
    assert worst <= TOL
    assert float(((V - out["fp64"]).abs() / rms_all).max().item()) <= TOL       # and what precision='auto' returned


def test_config5_pipeline_three_snapshots():
    """BASELINE config 5 shape (HERA-350 x 1024 ch + Tsys noise + windowed delay transform) for 3 of the 1000
    snapshots with a 20k-source catalogue: size-independent properties of the full pipeline."""
    from prisim_b200 import synthetic as S
    from prisim_b200.delay_spectrum import windowing
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    cfg = S.config5(nsnap=3, nsrc=20000)
    ia = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", pointing_coords="hadec", A_eff=cfg["A_eff"], eff_Q=cfg["eff_Q"], device=0, noise_seed=11)
    for j in range(3):
        ia.observe(SimpleTime(2451545.0 + j * 1e-4, j * 10.7 / 240.0), cfg["Tsysinfo"], NP.ones(1024), cfg["pointing_hadec"], cfg["skymodel"], cfg["t_acc"])
    ia.generate_noise(); ia.add_noise()
    window = 1024 * windowing(1024, "bhw", area_normalize=True)
    ia.delay_transform(pad=1.0, freq_wts=window, verbose=False)
    V, N, R = ia._skyvis, ia._noise, ia._rms
    # rms follows the radiometer equation; noise is white with that rms; vis = sky + noise
    Tsys = O.system_temperature(cfg["Tsysinfo"], cfg["channels"], 1)[0]
    rms_o = O.thermal_noise_rms(Tsys[None, :, None], cfg["A_eff"], cfg["eff_Q"], [cfg["t_acc"]], cfg["channels"][1] - cfg["channels"][0])[0, :, 0]
    assert NP.allclose(R[0][5].cpu().numpy(), rms_o, rtol=1e-12)
    z = (N[1] / (R[1] / NP.sqrt(2.0)))
    assert abs(z.real.std().item() - 1) < 2e-3 and abs(z.imag.std().item() - 1) < 2e-3 and abs(z.mean().abs().item()) < 1e-3
    assert torch.equal(ia._vis[2], V[2] + N[2])
    # linearity of the delay transform and Parseval with the window
    L = ia._lag
    assert (L["vis"][0] - L["skyvis"][0] - L["noise"][0]).abs().max().item() <= 1e-9 * L["vis"][0].abs().max().item()
    df = cfg["channels"][1] - cfg["channels"][0]
    w = torch.as_tensor(window).cuda()
    lhs = L["skyvis"][1].abs().pow(2).sum(dim=1) / (1024 * df ** 2)
    rhs = (V[1] * w).abs().pow(2).sum(dim=1)
    assert ((lhs - rhs).abs() / rhs).max().item() < 1e-10
    # bounded-memory streaming gives the same products, snapshot by snapshot, and frees the device memory
    ib = InterferometerArray(cfg["labels"], cfg["baselines"], cfg["channels"], telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", pointing_coords="hadec", A_eff=cfg["A_eff"], eff_Q=cfg["eff_Q"], device=0, noise_seed=11)
    seen = {}
    def sink(j, prod):
        seen[j] = {k: v.clone() for k, v in prod.items()}
    for j in range(3):
        ib.observe(SimpleTime(2451545.0 + j * 1e-4, j * 10.7 / 240.0), cfg["Tsysinfo"], NP.ones(1024), cfg["pointing_hadec"], cfg["skymodel"], cfg["t_acc"])
        if j == 1:
            assert ib.drain(sink, noise=True, delay_transform={"pad": 1.0, "freq_wts": window}) == 2
    assert ib.drain(sink, noise=True, delay_transform={"pad": 1.0, "freq_wts": window}) == 1 and len(ib._skyvis) == 0 and ib.n_acc == 3
    for j in range(3):
        assert torch.equal(seen[j]["skyvis_freq"], V[j]) and torch.equal(seen[j]["vis_noise_freq"], N[j])
        assert torch.equal(seen[j]["vis_freq"], ia._vis[j]) and torch.equal(seen[j]["vis_lag"], L["vis"][j])
    # foreground power is confined within the horizon delay limits (+ window main lobe) on a long baseline
    b = 61074
    blen = NP.linalg.norm(cfg["baselines"][b])
    lag_power = L["skyvis"][0][b].abs().pow(2).cpu().numpy()
    inside = NP.abs(ia.lags) <= blen / 299792458.0 + 4.0 / (1024 * df)
    assert lag_power[inside].sum() > 0.999 * lag_power.sum()


def test_gridded_healpix_beam_vs_oracle():
    """External HEALPix beam gathered on the device (scripts/run_prisim.py:1897-1908) through observe and through
    HealpixBeam.pbeam; config 3's 'Airy sampled on a HEALPix grid' beam variant at nside 32."""
    from prisim_b200 import synthetic as S
    from prisim_b200 import primary_beams as PB
    from prisim_b200.interferometry import InterferometerArray, SimpleTime
    nside = 32
    ra, dec = S.healpix_ring_centers(nside)                            # beam frame: pole = zenith, longitude = azimuth
    pix_altaz = NP.stack((dec, ra), 1)
    bf = NP.linspace(140e6, 160e6, 6)
    up = pix_altaz[:, 0] > 0
    beam = NP.full((ra.size, bf.size), 1e-7)
    beam[up] = NP.maximum(O.airy_disk_pattern(14.0, pix_altaz[up], bf, power=True), 1e-7)
    cfg = S.config1(nsrc=1500, nchan=96, nsnap=1)
    chans = cfg["channels"]
    rng = NP.random.default_rng(5)
    altaz = NP.stack((rng.uniform(0.0, 90.0, 300), rng.uniform(0, 360, 300)), 1)
    for dtype, tol in ((torch.float64, 1e-11), (torch.float32, 3e-6)):
        hb = PB.HealpixBeam(beam, bf, chans, spec_interp="cubic", dtype=dtype, device=0)
        ref = O.external_beam_table(beam, bf, altaz, chans, kind="cubic")
        assert NP.abs(hb.pbeam(altaz) - ref).max() <= tol
    hb_a = PB.HealpixBeam(beam, bf, chans, chromatic=False, select_freq=151e6, dtype=torch.float64, device=0)
    assert NP.abs(hb_a.pbeam(altaz) - O.external_beam_table(beam, bf, altaz, chans, chromatic=False, select_freq=151e6)).max() <= 1e-11
    # through observe: same visibilities as the reference's route (beam table handed over in roi_info)
    hb = PB.HealpixBeam(beam, bf, chans, spec_interp="cubic", dtype=torch.float32, device=0)
    ia = InterferometerArray(cfg["labels"], cfg["baselines"], chans, telescope=cfg["telescope"], latitude=cfg["latitude"],
                             skycoords="radec", pointing_coords="hadec", device=0)
    ia.observe(SimpleTime(2451545.0, 33.0), {"Tnet": 200.0}, NP.ones(96), cfg["pointing_hadec"], cfg["skymodel"], cfg["t_acc"],
               pb_info={"external_beam": hb})
    sky = cfg["skymodel"]; sp = sky.spec_parms
    hadec = NP.stack((33.0 - sky.location[:, 0], sky.location[:, 1]), 1)
    src_altaz = O.hadec2altaz(hadec, cfg["latitude"])
    m2 = O.roi_select(src_altaz)
    pbeam = O.external_beam_table(beam, bf, src_altaz[m2], chans, kind="cubic")
    Vo, _ = O.observe_snapshot(cfg["baselines"], chans, hadec, "hadec", cfg["latitude"], cfg["pointing_hadec"], "hadec", cfg["telescope"],
                               sp["flux-scale"], sp["power-law-index"], sp["freq-ref"], roi_info={"ind": m2, "pbeam": pbeam})
    assert NP.array_equal(ia.obs_catalog_indices[0], m2)
    assert rel_err(ia.skyvis_freq[:, :, 0], Vo) <= TOL
    with pytest.raises(ValueError):
        PB.HealpixBeam(beam[:-1], bf, chans)


def test_peer_gather_buffer_single_rank():
    """sharding.PeerGatherBuffer on a one-rank NCCL group: the library-allocated, IPC-exported buffer is a valid
    output target for the phase-sum kernel (the 2- and 8-GPU runs of bench.py use the same path across ranks)."""
    import socket
    import torch.distributed as dist
    from prisim_b200 import engine
    from prisim_b200.sharding import PeerGatherBuffer
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    dist.init_process_group("nccl", init_method="tcp://127.0.0.1:{0}".format(port), world_size=1, rank=0,
                            device_id=torch.device("cuda:0"))
    try:
        nbl, nchan, nsrc = 70, 130, 90
        g = PeerGatherBuffer((nbl, nchan), 0, dst=0)
        assert g.mode == "peer" and g.full.shape == (1, nbl, nchan)
        rng = NP.random.default_rng(1)
        bl = rng.normal(0, 100, (nbl, 3)); freqs = 150e6 + NP.arange(nchan) * 1e5
        altaz = NP.stack((NP.degrees(NP.arcsin(rng.uniform(0, 1, nsrc))), rng.uniform(0, 360, nsrc)), 1)
        dircos, _ = engine.sky_cull(altaz, "altaz")
        amp = engine.dense_to_amp_table(torch.rand((nsrc, nchan), device="cuda"))
        engine.skyvis(dircos, amp, nsrc, bl, (0, 0, 1.0), freqs, out=g.local)
        ref = engine.skyvis(dircos, amp, nsrc, bl, (0, 0, 1.0), freqs)
        g.wait()
        assert torch.equal(g.full[0], ref)
        g.close()
    finally:
        dist.destroy_process_group()
