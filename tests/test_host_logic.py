"""CPU tests: host-side helpers, the C-ABI library surface, and loud failure without a GPU."""
import ctypes
import os
import re

import numpy as NP
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ensure_built():
    import __graft_entry__ as G
    if not os.path.exists(os.path.join(ROOT, "prisim_b200", "libprisim_b200.so")):
        G.build()


def test_library_loads_and_exports_every_declared_symbol():
    _ensure_built()
    from prisim_b200 import _lib
    header = open(os.path.join(ROOT, "include", "prisim_b200.h")).read()
    declared = set(re.findall(r"\b(pb200_[a-z_0-9]+)\s*\(", header))
    declared -= {"pb200_ctx", "pb200_beam_desc", "pb200_spectrum_desc"}
    assert len(declared) >= 13
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "symbol {0} declared in the header is not exported".format(name)
        assert name in _lib.SYMBOLS, "symbol {0} has no ctypes prototype".format(name)
    assert set(_lib.SYMBOLS) == declared
    assert lib.pb200_version() == 200          # round 2: pb200_skyvis gained nsrc_bright + vis_row_stride, pb200_noise bl_step


def test_struct_layouts_match_header_sizes():
    from prisim_b200 import _lib
    # pb200_beam_desc: 4 int32 + 11 doubles(+3+3 arrays) ... computed from the C declaration order
    assert ctypes.sizeof(_lib.BeamDesc) == 4 * 4 + 8 * (1 + 3 + 3 + 4) + 2 * 4 + 8 * 3 + 8 * 3 + 2 * 4 + 4 * 8
    assert ctypes.sizeof(_lib.SpectrumDesc) == 5 * 8


def test_host_only_entry_points():
    _ensure_built()
    from prisim_b200 import _lib, engine
    lib = _lib.load()
    assert lib.pb200_nsrc_pad(1) == 32 and lib.pb200_nsrc_pad(32) == 32 and lib.pb200_nsrc_pad(33) == 64
    assert lib.pb200_amp_bytes(33, 129) == 2 * 64 * 128 * 4
    assert engine.delay_nout(1024, 1.0, True) == 1024
    assert engine.delay_nout(1024, 0.0, True) == 1024
    assert engine.delay_nout(64, 0.5, True) == 64          # ceil(96 / 1.5)
    assert engine.delay_nout(64, 1.0, False) == 128
    assert engine.delay_nout(10, 0.25, True) == 10         # ceil(12 / 1.25)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU failure mode")
def test_product_path_fails_loudly_without_gpu():
    _ensure_built()
    from prisim_b200 import _lib, engine
    with pytest.raises(_lib.PB200Error):
        _lib.get_context(0)
    with pytest.raises(_lib.PB200Error):
        engine.sky_cull(NP.zeros((4, 2)), "altaz")


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "prisim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f


def test_hexagon_and_baselines():
    from prisim_b200.interferometry import baseline_generator, hexagon_generator, orient_and_sort_baselines, uniq_baselines
    xy, labels = hexagon_generator(14.6, n_side=3)
    assert xy.shape == (19, 2) and len(labels) == 19
    assert NP.allclose(xy.mean(axis=0), 0.0, atol=1e-12)
    d = NP.sqrt(((xy[:, None, :] - xy[None, :, :]) ** 2).sum(-1))
    assert NP.isclose(d[d > 0].min(), 14.6)
    xy2, _ = hexagon_generator(14.6, n_total=19)
    assert NP.allclose(xy, xy2)
    with pytest.raises(ValueError):
        hexagon_generator(14.6, n_total=20)
    assert hexagon_generator(14.6, n_side=11)[0].shape == (331, 2)
    ant = NP.hstack((xy, NP.zeros((19, 1))))
    bl, lab, ids = baseline_generator(ant, ant_label=NP.arange(19).astype(str))
    assert bl.shape == (171, 3)
    # b = r_j - r_i, i outer loop, j > i  (interferometry.py:1355-1358)
    assert NP.allclose(bl[0], ant[1] - ant[0]) and NP.allclose(bl[17], ant[18] - ant[0]) and NP.allclose(bl[18], ant[2] - ant[1])
    assert lab["A2"][0] == "1" and lab["A1"][0] == "0"
    bls, labs, order = orient_and_sort_baselines(bl, lab)
    blo = NP.degrees(NP.angle(bls[:, 0] + 1j * bls[:, 1]))
    assert NP.all((blo > -67.5 - 1e-9) & (blo <= 112.5 + 1e-9))
    assert NP.all(NP.diff(NP.sqrt((bls ** 2).sum(1))) >= -1e-12)
    ub, first, counts, occ = uniq_baselines(bls)
    assert ub.shape[0] == 30 and counts.sum() == 171                     # HERA-19: 30 unique baselines
    bla, _, _ = baseline_generator(ant, auto=True)
    assert bla.shape == (190, 3)
    blc, _, _ = baseline_generator(ant, conjugate=True)
    assert blc.shape == (342, 3)


def test_synthetic_configs_shapes():
    from prisim_b200 import synthetic as S
    c1 = S.config1()
    assert c1["baselines"].shape == (171, 3) and c1["channels"].size == 128 and c1["skymodel"].location.shape == (1000, 2)
    assert NP.isclose(c1["channels"][64], 150e6) and NP.isclose(c1["channels"][1] - c1["channels"][0], 100e3)
    c2 = S.config2(nsrc=100)
    assert c2["ant"].shape == (350, 3) and c2["baselines"].shape == (61075, 3) and c2["channels"].size == 1024
    assert c2["skymodel"].location[:, 1].max() <= 30.0
    ra, dec = S.healpix_ring_centers(8)
    assert ra.size == 768 and NP.all(NP.diff(dec) <= 1e-12)              # RING order: Dec non-increasing
    v = NP.stack((NP.cos(NP.radians(dec)) * NP.cos(NP.radians(ra)), NP.cos(NP.radians(dec)) * NP.sin(NP.radians(ra)), NP.sin(NP.radians(dec))), 1)
    assert NP.abs(v.mean(axis=0)).max() < 1e-12
    c4 = S.config4(ntiles=16, nsrc=50, nchan=8)
    assert c4["baselines"].shape == (120, 3) and NP.allclose(NP.mod(c4["pb_info"]["delays"] / 435e-12 + 0.5, 1.0), 0.5)


def test_skymodel_container():
    from prisim_b200.skymodel import SkyModel
    parms = {"location": [[10.0, -30.0], [20.0, 5.0]], "spec_type": "func", "frequency": [150e6],
             "spec_parms": {"name": NP.repeat("power-law", 2), "power-law-index": [-0.8, -0.7], "freq-ref": [150e6, 150e6],
                            "flux-scale": [1.0, 2.0]}}
    sm = SkyModel(init_parms=parms)
    sp = sm.generate_spectrum(frequency=[150e6, 300e6])
    assert NP.allclose(sp, [[1.0, 2 ** -0.8], [2.0, 2 * 2 ** -0.7]])
    assert NP.allclose(sm.generate_spectrum(ind=[1], frequency=[300e6]), [[2 * 2 ** -0.7]])
    sub = sm.subset([1])
    assert sub.location.shape == (1, 2) and sub.spec_parms["flux-scale"][0] == 2.0
    tab = SkyModel(init_parms={"location": [[0.0, 0.0]], "spec_type": "spectrum", "frequency": [100e6, 200e6], "spectrum": [[1.0, 3.0]]})
    assert NP.allclose(tab.generate_spectrum(frequency=[150e6], interp_method="linear"), [[2.0]])
    with pytest.raises(TypeError):
        SkyModel(init_parms=None)


def test_shard_bounds():
    from prisim_b200.sharding import shard_bounds, shard_slice
    b = shard_bounds(61075, 8)
    assert b[0] == 0 and b[-1] == 61075 and NP.all(NP.diff(b) >= 7634) and NP.all(NP.diff(b) <= 7635)
    assert shard_slice(10, 4, 3) == slice(8, 10)
    assert list(shard_bounds(3, 4)) == [0, 1, 2, 3, 3]


def test_interferometer_array_validation_without_gpu():
    from prisim_b200.interferometry import InterferometerArray
    bl = NP.zeros((3, 3)); ch = 150e6 + NP.arange(4) * 1e5
    with pytest.raises(ValueError):
        InterferometerArray(["a", "b"], bl, ch)
    with pytest.raises(TypeError):
        InterferometerArray("abc", bl, ch)
    with pytest.raises(ValueError):
        InterferometerArray(["a", "b", "c"], bl, ch, skycoords="galactic")
    with pytest.raises(ValueError):
        InterferometerArray(["a", "b", "c"], bl, ch, eff_Q=[1.5, 0.2, 0.3])
    ia = InterferometerArray(["a", "b", "c"], bl[:, :2], ch / 1e6, freq_scale="MHz", A_eff=[1.0, 2.0, 3.0])
    assert ia.baselines.shape == (3, 3) and NP.allclose(ia.channels, ch) and ia.A_eff.shape == (3, 4)
    assert NP.isclose(ia.freq_resolution, 1e5) and ia.n_acc == 0 and ia.skyvis_freq is None


def test_subband_windows_and_window_helpers_match_oracle():
    """Host-only parts of multi_window_delay_transform (interferometry.py:8199-8265): no GPU needed."""
    from prisim_b200.interferometry import InterferometerArray
    from prisim_b200.delay_spectrum import window_N2width, windowing
    from oracle import prisim_oracle as O
    ch = 150e6 + (NP.arange(96) - 48) * 1e5
    ia = InterferometerArray(["a", "b"], NP.zeros((2, 3)), ch)
    for shape in ("rect", "bhw", "bnw", "BHW"):
        assert abs(window_N2width(shape=shape) - O.window_N2width(shape)) < 1e-15
        assert NP.array_equal(windowing(37, shape=shape.lower()), O.windowing(37, shape=shape.lower()))
        for bw, fc in ((1.0e6, None), ([0.8e6, 2.5e6, 1.1e6], [ch[70], ch[5], ch[40]]), (3.0e6, [ch[2], ch[93]]), ([1e6, 2e6], ch[50])):
            w = ia.subband_windows(bw, freq_center=fc, shape=shape)
            assert NP.array_equal(w, O.multi_window_weights(ch, bw, fc, shape))
            assert w.shape[1] == ch.size and NP.all(w >= 0.0) and NP.all(w.max(axis=1) <= 1.0 + 1e-15)
    w = ia.subband_windows([1e6, 1e6], freq_center=[ch[80], ch[10]], shape="bhw")
    assert NP.argmax(w[0]) < NP.argmax(w[1])                                     # ordered by centre channel (:8247-8251)
    w = ia.subband_windows(4.0e6, freq_center=ch[1], shape="rect")               # clipped at the band edge, not wrapped
    assert w[0, 0] == 1.0 and w[0, -1] == 0.0
    for bad in (dict(bw_eff=0.0), dict(bw_eff=[1e6, 2e6], freq_center=[ch[3], ch[4], ch[5]]), dict(bw_eff=1e6, freq_center=ch[-1]),
                dict(bw_eff=1e6, shape="hann")):
        with pytest.raises(ValueError):
            ia.subband_windows(**bad)
    with pytest.raises(TypeError):
        ia.subband_windows("1 MHz")
    with pytest.raises(TypeError):
        ia.subband_windows(1e6, shape=3)


def test_roi_parameters_host_paths_without_gpu():
    """ROI_parameters.append_settings (interferometry.py:4221-4617): the branches that need no beam evaluation."""
    from prisim_b200.interferometry import ROI_parameters
    from prisim_b200.skymodel import SkyModel
    n = 20
    sky = SkyModel(init_parms={"location": NP.stack((NP.linspace(0, 340, n), NP.linspace(-80, 20, n)), axis=1), "coords": "hadec",
                               "spec_type": "func", "frequency": [150e6],
                               "spec_parms": {"name": NP.repeat("power-law", n), "power-law-index": NP.zeros(n), "freq-ref": NP.full(n, 150e6),
                                              "flux-scale": NP.ones(n)}})
    freq = 0.15 + NP.arange(8) * 1e-4                                            # GHz (the reference's default freq_scale)
    tel = {"id": "hera", "shape": "dish", "size": 14.0, "orientation": NP.asarray([90.0, 270.0]), "ocoords": "altaz", "latitude": -30.7}
    roi = ROI_parameters()
    pb = NP.random.default_rng(0).uniform(0, 1, (3, 8))
    roi.append_settings(sky, freq, roi_info={"ind": NP.asarray([1, 5, 9]), "pbeam": pb, "radius": 90.0}, telescope=tel)
    assert NP.allclose(roi.freq, freq * 1e9) and roi.freq_scale == "Hz"
    assert roi.info["pbeam"][0].dtype == NP.float32 and NP.array_equal(roi.info["pbeam"][0], pb.astype(NP.float32))
    assert NP.array_equal(roi.info["ind"][0], [1, 5, 9]) and roi.info["radius"] == [90.0] and roi.pinfo == []
    roi.append_settings(None, freq, telescope=tel)
    assert roi.info["ind"][1].size == 0 and roi.info["pbeam"][1].size == 0 and roi.pinfo == [None]
    with pytest.raises(ValueError):
        roi.append_settings(sky, freq, roi_info=None)
    with pytest.raises(ValueError):
        roi.append_settings(sky, freq, roi_info={"ind": NP.arange(4), "pbeam": NP.ones((3, 8))})
    with pytest.raises(ValueError):
        roi.append_settings(sky, freq, roi_info={"ind": NP.arange(3), "pbeam": NP.ones((3, 7))})
    with pytest.raises(TypeError):
        roi.append_settings("sky", freq, roi_info={"radius": 10.0})
    with pytest.raises(ValueError):                                              # ROI by radius needs the pointing info for the beam
        roi.append_settings(sky, freq, pinfo=None, roi_info={"radius": 30.0, "center": None})
    with pytest.raises(TypeError):
        ROI_parameters().append_settings(sky, freq, roi_info={"radius": 10.0}, telescope=None)
    with pytest.raises(IOError):                                                 # a missing file (FITS interchange itself: test below)
        ROI_parameters(init_file="no_such_roi.fits")


def test_gradient_and_duplicate_argument_errors_without_gpu():
    from prisim_b200.interferometry import InterferometerArray
    ia = InterferometerArray(["a", "b"], NP.zeros((2, 3)), 150e6 + NP.arange(4) * 1e5)
    assert ia.gradient == {} and ia.gradient_mode is None
    with pytest.raises(AttributeError):                                          # interferometry.py:6765-6769
        ia.apply_gradients(perturbations={"baseline": NP.zeros((3, 2))})
    with pytest.raises(TypeError):                                               # :6849-6850
        ia.duplicate_measurements(blgroups=[("a", "b")])
    ia.duplicate_measurements(blgroups={"a": ["a"], "b": ["b"]})                 # nothing to expand: no-op (:6852-6857)
    assert ia.baselines.shape[0] == 2


def test_uniq_baselines_matches_reference_and_groups_feed_duplicate_measurements():
    """uniq_baselines (interferometry.py:1373-1463) against the reference's own run, shim and oracle; baseline_groups builds
    the blgroups / reversemap dictionaries of interferometry.py:1999-2007."""
    from prisim_b200.interferometry import uniq_baselines, baseline_groups
    from oracle import prisim_oracle as O
    g = NP.load(os.path.join(ROOT, "tests", "golden", "uniq_baselines.npz"))
    for key, red in (("all", None), ("red", True), ("nonred", False)):
        for fn in (uniq_baselines, O.uniq_baselines):
            sel, ind, cnt, occ = fn(g["bl"], redundant=red)
            assert NP.array_equal(sel, g["sel_" + key]) and NP.array_equal(ind, g["ind_" + key]) and NP.array_equal(cnt, g["cnt_" + key])
            assert [len(o) for o in occ] == g["occ_len_" + key].tolist()
            assert NP.array_equal(NP.concatenate([NP.asarray(o, dtype=int) for o in occ]), g["occ_flat_" + key])
    with pytest.raises(TypeError):
        uniq_baselines(g["bl"].tolist())
    with pytest.raises(TypeError):
        uniq_baselines(g["bl"], redundant="yes")
    assert uniq_baselines(g["bl"][:, :2])[0].shape[1] == 3                       # 2-column input is zero-filled (:1425-1426)
    # groups: every baseline in exactly one group, keys are group members, reverse map points back to the key
    dt = [("A2", "U4"), ("A1", "U4")]
    labels = NP.asarray([("j{0}".format(i), "i{0}".format(i)) for i in range(g["bl"].shape[0])], dtype=dt)
    ulab, ubl, info = baseline_groups(labels, g["bl"])
    assert ubl.shape == (69, 3) and len(info["groups"]) == 69 and len(info["reversemap"]) == 210
    members = [tuple(m) for v in info["groups"].values() for m in v.tolist()]
    assert sorted(members) == sorted(tuple(l) for l in labels.tolist())
    for key, v in info["groups"].items():
        assert key in [tuple(m) for m in v.tolist()]
        for m in v.tolist():
            assert tuple(info["reversemap"][tuple(m)][0].tolist()) == key
    assert [tuple(l) for l in ulab.tolist()] == list(info["groups"].keys())


def test_layouts_match_reference_golden():
    """hexagon_generator / baseline_generator (interferometry.py:857-989, :1184-1370) against the reference's own run."""
    from prisim_b200.interferometry import hexagon_generator, baseline_generator
    g = NP.load(os.path.join(ROOT, "tests", "golden", "layouts.npz"))
    for tag, kw in (("side3", dict(n_side=3)), ("side11", dict(n_side=11)), ("side4_rot", dict(n_side=4, orientation=30.0, center=NP.asarray([[5.0, -3.0]])))):
        xy, lab = hexagon_generator(14.6, **kw)
        assert NP.array_equal(xy, g["hex_" + tag]) and list(lab) == g["hexlab_" + tag].tolist()
    assert NP.array_equal(hexagon_generator(14.6, n_total=331)[0], g["hex_side11"])     # the n_total route (a 1-element array index in the reference)
    with pytest.raises(ValueError):
        hexagon_generator(14.6, n_total=330)
    ant = NP.hstack((g["hex_side3"], 0.1 * NP.arange(19).reshape(-1, 1)))
    for tag, kw in (("plain", {}), ("auto", dict(auto=True)), ("conj", dict(conjugate=True))):
        bl, lab, ids = baseline_generator(ant, **kw)
        assert NP.array_equal(bl, g["bl_" + tag])
        assert [[str(a), str(b)] for a, b in lab.tolist()] == g["bllab_" + tag].tolist()
        assert [list(r) for r in ids.tolist()] == g["blid_" + tag].tolist()


def test_thermal_noise_rms_matches_reference_golden():
    """Module-level thermalNoiseRMS (interferometry.py:89-230) against the reference's own run."""
    from prisim_b200.interferometry import thermalNoiseRMS
    g = NP.load(os.path.join(ROOT, "tests", "golden", "thermal_rms.npz"))
    Ts = g["Tsys"]
    nb, nc, nt = Ts.shape
    assert NP.allclose(thermalNoiseRMS(12.5, 1e5, 10.7, Ts, nbl=nb, nchan=nc, ntimes=nt, flux_unit="Jy", eff_Q=0.9), g["jy_full"], rtol=1e-14)
    assert NP.allclose(thermalNoiseRMS(12.5, 1e5, 10.7, Ts, nbl=nb, nchan=nc, ntimes=nt, flux_unit="K", eff_Q=0.9), g["k_full"], rtol=1e-14)
    r = thermalNoiseRMS(NP.full((1, nc, 1), 20.0), 1e5, 5.0, Ts[:1, :, :1], nbl=nb, nchan=nc, ntimes=nt, eff_Q=NP.linspace(0.5, 1.0, nb).reshape(nb, 1, 1))
    assert r.shape == g["jy_chan"].shape and NP.allclose(r, g["jy_chan"], rtol=1e-14)
    assert NP.allclose(thermalNoiseRMS(100.0, 2e5, 1.0, 250.0), g["jy_scalar"], rtol=1e-14)
    with pytest.raises(IndexError):
        thermalNoiseRMS(1.0, 1e5, 1.0, NP.ones((2, 2, 2)), nbl=3, nchan=2, ntimes=2)
    with pytest.raises(ValueError):
        thermalNoiseRMS(1.0, 1e5, 1.0, -5.0)
    with pytest.raises(ValueError):
        thermalNoiseRMS(1.0, 1e5, 1.0, 5.0, flux_unit="mJy")
    with pytest.raises(TypeError):
        thermalNoiseRMS(1.0, [1e5], 1.0, 5.0)
    with pytest.raises(TypeError):
        thermalNoiseRMS(1.0, 1e5, 1.0, 5.0, nbl=2.0)


def test_baseline_group_lookups_match_reference_golden():
    """getBaselineGroupKeys / getBaselinesInGroups (interferometry.py:2017-2165) against the reference's own run."""
    from prisim_b200.interferometry import getBaselineGroupKeys, getBaselinesInGroups, baseline_groups
    g = NP.load(os.path.join(ROOT, "tests", "golden", "uniq_baselines.npz"))
    n_ant = 21
    ii, jj = NP.triu_indices(n_ant, k=1)
    labs = NP.asarray([(str(j), str(i)) for i, j in zip(ii, jj)], dtype=[("A2", "U3"), ("A1", "U3")])
    _, _, info = baseline_groups(labs, g["bl"])
    query = [tuple(q) for q in g["query"].tolist()]
    keys, flipped = getBaselineGroupKeys(query, info["reversemap"])
    members, flipped2 = getBaselinesInGroups(query, info["reversemap"], info["groups"])
    assert flipped == flipped2
    for k, f, m, kr, fr, cr, mr in zip(keys, flipped, members, g["keys"].tolist(), g["flipped"].tolist(), g["member_counts"].tolist(), g["member_first"].tolist()):
        if fr < 0:
            assert k is None and f is None and m is None
        else:
            assert list(k) == kr and int(f) == fr and m.size == cr and list(m[0].tolist()) == mr
    with pytest.raises(TypeError):
        getBaselineGroupKeys(query, [("1", "0")])
    with pytest.raises(TypeError):
        getBaselinesInGroups(query, info["reversemap"], None)


def test_fits_min_roundtrip_and_card_format(tmp_path):
    """prisim_b200.fits_min: the fixed-format FITS subset ROI_parameters.save needs (primary header + IMAGE extensions)."""
    from prisim_b200 import fits_min as F
    rng = NP.random.default_rng(0)
    arrays = {"F8": rng.normal(size=(5, 7)), "I8": rng.integers(-5, 10 ** 12, 13), "F4": rng.normal(size=(3, 4, 2)).astype(NP.float32),
              "I4": rng.integers(0, 1000, (2, 3)).astype(NP.int32), "U1": rng.integers(0, 255, 9).astype(NP.uint8), "EMPTYDIM": NP.zeros((0, 4))}
    fn = str(tmp_path / "t.fits")
    hdr = {"n_obs": (2, "Number of observations"), "element_shape": ("dish", "Antenna element shape"), "element_size": (14.0, "m"),
           "latitude": -30.7224, "flag": True, "quote": "it's", "tiny": 1.5e-300}
    F.write(fn, hdr, [(k, v, {"delayerr": (0.0, "Jitter in delays [s]")} if k == "I8" else None) for k, v in arrays.items()])
    raw = open(fn, "rb").read()
    assert len(raw) % 2880 == 0 and raw[:30] == b"SIMPLE  =                    T"
    first = raw[:2880].decode("ascii")
    cards = [first[i:i + 80] for i in range(0, 2880, 80)]
    assert any(c.startswith("HIERARCH element_shape = 'dish    '") for c in cards) and any(c.startswith("END") for c in cards)
    hdus = F.read(fn)
    h0 = hdus[0][0]
    assert F.header_get(h0, "N_OBS") == 2 and F.header_get(h0, "element_size") == 14.0 and F.header_get(h0, "LATITUDE") == -30.7224
    assert F.header_get(h0, "flag") is True and F.header_get(h0, "quote") == "it's" and F.header_get(h0, "tiny") == 1.5e-300
    for k, v in arrays.items():
        got = F.getdata(fn, k.lower())
        assert got.dtype == v.dtype and got.shape == v.shape and NP.array_equal(got, v), k
    assert F.header_get(hdus[2][0], "delayerr") == 0.0
    with pytest.raises(KeyError):
        F.getdata(fn, "nope")
    with pytest.raises(IOError):
        F.write(fn, {}, [])
    with pytest.raises(TypeError):
        F.write(str(tmp_path / "c.fits"), {}, [("C", NP.zeros(3, dtype=NP.complex128), None)])


def test_roi_parameters_save_and_init_file_roundtrip(tmp_path):
    """ROI_parameters.save / ROI_parameters(init_file=...) with the reference's FITS layout (interferometry.py:4621-4723, :4080-4205)."""
    from prisim_b200 import fits_min as F
    from prisim_b200.interferometry import ROI_parameters
    from prisim_b200.skymodel import SkyModel
    rng = NP.random.default_rng(1)
    nsrc, nchan = 40, 6
    parms = {"location": NP.stack((rng.uniform(0, 360, nsrc), rng.uniform(-90, 90, nsrc)), 1), "coords": "hadec", "spec_type": "func", "frequency": [150e6],
             "spec_parms": {"name": NP.repeat("power-law", nsrc), "power-law-index": NP.zeros(nsrc), "freq-ref": NP.full(nsrc, 150e6), "flux-scale": NP.ones(nsrc)}}
    sky = SkyModel(init_parms=parms)
    tel = {"id": "mwa", "shape": "dipole", "size": 0.74, "orientation": NP.asarray([1.0, 0.0, 0.0]), "ocoords": "dircos", "groundplane": 0.3,
           "ground_modify": {"scale": 1.5, "max": 4.0}, "latitude": -26.701, "longitude": 116.67, "altitude": 377.8,
           "element_locs": rng.normal(size=(16, 3))}
    freq = 0.15 + 1e-4 * NP.arange(nchan)
    roi = ROI_parameters()
    inds = [NP.asarray([3, 17, 25]), NP.asarray([], dtype=int), NP.asarray([0, 1, 2, 39])]
    for j, ind in enumerate(inds):
        pb = rng.uniform(0, 1, (ind.size, nchan))
        roi.append_settings(sky, freq, pinfo={}, roi_info={"ind": ind, "pbeam": pb if ind.size else None}, telescope=tel, freq_scale="GHz")
    roi.pinfo = [{"delays": rng.uniform(0, 1e-8, 16), "delayerr": None, "pointing_center": NP.asarray([[60.0, 120.0]]), "pointing_coords": "altaz"},
                 {"delays": rng.uniform(0, 1e-8, 16), "delayerr": 2e-10}, {"pointing_center": NP.asarray([[0.1, 0.2, 0.97]]), "pointing_coords": "dircos"}]
    base = str(tmp_path / "roiinfo")
    roi.save(base, verbose=False)
    with pytest.raises(IOError):
        roi.save(base, verbose=False)
    roi.save(base, overwrite=True, verbose=False)
    # what scripts/run_prisim.py:1959-1961 reads per (chunk, snapshot)
    assert NP.array_equal(F.getdata(base + ".fits", "IND_2"), inds[2])
    assert NP.array_equal(F.getdata(base + ".fits", "PB_0"), roi.info["pbeam"][0])
    back = ROI_parameters(init_file=base + ".fits")
    assert NP.allclose(back.freq, freq * 1e9) and len(back.info["ind"]) == 3
    for j in range(3):
        assert NP.array_equal(back.info["ind"][j], inds[j]) and NP.array_equal(back.info["pbeam"][j], roi.info["pbeam"][j])
    t = back.telescope
    assert t["id"] == "mwa" and t["shape"] == "dipole" and t["size"] == 0.74 and t["ocoords"] == "dircos" and t["groundplane"] == 0.3
    assert t["ground_modify"] == {"scale": 1.5, "max": 4.0} and t["latitude"] == -26.701 and t["longitude"] == 116.67 and t["altitude"] == 377.8
    assert NP.array_equal(t["orientation"], NP.asarray([[1.0, 0.0, 0.0]])) and NP.array_equal(t["element_locs"], tel["element_locs"])
    assert NP.array_equal(back.pinfo[0]["delays"], roi.pinfo[0]["delays"]) and back.pinfo[0]["delayerr"] is None
    assert back.pinfo[0]["pointing_coords"] == "altaz" and NP.array_equal(back.pinfo[0]["pointing_center"], roi.pinfo[0]["pointing_center"])
    assert back.pinfo[1]["delayerr"] == 2e-10 and "pointing_center" not in back.pinfo[1]
    assert back.pinfo[2]["pointing_coords"] == "dircos" and "delays" not in back.pinfo[2]


def test_fixed_point_anchor_phase_identity():
    """The quarter-block kernel (csrc/skyvis.cu, PB_Q3_INT) stages, per (source, baseline), the anchor phase at the CTA's
    first channel and its step per 32-channel block as 32-bit fixed-point turn fractions; a thread's phase is the wrapping
    integer sum X0 + block * D.  Numpy restatement of that arithmetic against the directly reduced phase: the difference
    stays below (1 + 15) / 2 * 2^-32 turn (1.2e-8 rad, a tenth of an fp32 ulp of the MUFU argument)."""
    rng = NP.random.default_rng(77)
    tau = rng.uniform(-1.2e-6, 1.2e-6, 20000)                               # seconds: |b| <= 360 m
    df = 97656.25
    fcta0 = 100e6 + rng.integers(0, 8, tau.size) * 512 * df                  # first channel of a 4-slab CTA tile
    block = rng.integers(0, 16, tau.size)
    frac = lambda x: x - NP.rint(x)
    to_fix = lambda x: NP.rint(frac(x) * 4294967296.0).astype(NP.int64) & 0xFFFFFFFF
    X = (to_fix(tau * fcta0) + block * to_fix(tau * (32.0 * df))) & 0xFFFFFFFF
    turns = NP.where(X >= 2 ** 31, X - 2 ** 32, X) / 4294967296.0
    direct = frac(tau * (fcta0 + block * 32.0 * df))
    assert NP.abs(frac(turns - direct)).max() <= 8.5 * 2.0 ** -32
