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
    assert lib.pb200_version() == 100


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
    ub, first, counts = uniq_baselines(bls)
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
